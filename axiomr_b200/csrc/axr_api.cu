// libaxr_b200.so — context management and the C ABI declared in include/axr_b200.h.
// Host-side restatement of the bookkeeping in TiledPipeline::drawMesh (reference src/tiled_pipeline.cpp:143-322):
// uniforms, per-draw buffers sized from counts (instead of the fixed 32 MB arenas, include/tiled_pipeline.hpp:98-107),
// kernel sequencing on one CUDA stream (instead of ThreadPool futures, :183-248, :281-308).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/axr_b200.h"
#include "axr_kernels.cuh"
#include "axr_tangents.cuh"
#include "axr_obj.hpp"

using namespace axr;

namespace {

thread_local std::string g_create_error;

struct DeviceMesh {
	bool live = false;
	float4* pos = nullptr;
	float4* attr = nullptr;   // 48 B per vertex (three float4 planes, or records: AXR_ATTR_PLANES)
	uint4* idx4 = nullptr;    // AXR_IDX_PAD only
	unsigned* idx = nullptr;
	float4* sv[2] = {nullptr, nullptr};  // per-draw screen-space vertex records (16 B each), one per draw slot
	unsigned long long n_verts = 0, n_faces = 0;
	std::vector<unsigned long long> group_first;  // n_groups + 1; faces before group_first[0] belong to no group and are not drawn
	std::vector<std::string> group_names;         // axr_load_obj only
	std::vector<float> host_vertices;             // axr_load_obj only: the loader's arrays (AR::Vertex layout), for axr_mesh_read
	std::vector<uint32_t> host_indices;
	std::vector<Material> materials;              // host copy
	Material* d_materials = nullptr;
	unsigned long long* d_group_first = nullptr;
	bool materials_dirty = true;
	int bin_mode = BINS_LISTS;  // how the next draw of this mesh is issued, learnt from its last one: no triangle went through the bins
	                            // -> BINS_NONE, a few -> BINS_SCAN (both without the two bin kernels; a draw that turns out to need
	                            // them raises OVF_NEED_BINS and is re-issued with BINS_LISTS)
};

struct DeviceTexture {
	bool live = false;
	uchar4* data = nullptr;  // tiled order (axr_shaders.cuh: tex_offset_x / tex_offset_y)
	int w = 0, h = 0, tiles_x = 0;
};

// A shader functor compiled by its author and opened at run time (axr_load_shader_plugin, include/axr_shader_plugin.cuh)
struct ShaderPlugin {
	void* handle = nullptr;
	int (*launch)(const void*, const void*, const void*, const void*, int, int, int, unsigned, unsigned, void*) = nullptr;
	bool discards = false;
	unsigned textures = 0;
	std::string path;
};

struct PendingDraw {
	bool valid = false;
	int slot = 0;
	axr_mesh mesh = -1;
	float model[16];
	int redo_depth = 0;
	bool peel = false;
	bool bins = true;  // issued with the bin kernels
};

}  // namespace

struct axr_ctx {
	int device = 0;
	FrameParams fp{};
	int sampler = 0;
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	std::string error;

	// framebuffer (full-frame pitch; only the band rows are touched)
	unsigned* color = nullptr;
	float* depth = nullptr;
	unsigned* out_color = nullptr;  // where the resolve stores go (own buffers unless axr_set_output redirected them)
	float* out_depth = nullptr;

	// per-frame state
	float view_proj[16];
	float viewport[16];
	float cam_pos[3];
	int shader_kind = AXR_SHADER_FLAT;
	axr_shader_params shader_params{};
	float shader_user[8] = {};
	std::vector<ShaderPlugin> plugins;  // shader kind AXR_SHADER_PLUGIN_BASE + i

	// Raster state of one draw in flight. Two slots alternate so that the geometry stages of draw i+1 (vertex, setup, bins — on
	// geom_stream) overlap the tile / shading kernel of draw i (on the main stream): the two halves stress different parts of
	// the SM (latency-bound gathers vs FP32 issue), so running them side by side raises throughput.
	struct DrawSlot {
		unsigned long long* vis = nullptr;
		unsigned* tile_touched = nullptr;
		unsigned* tile_count = nullptr;
		unsigned* bin_start = nullptr;
		unsigned* items = nullptr;
		unsigned ref_cap = 0;
		TriRecord* records = nullptr;
		unsigned rec_cap = 0;
		unsigned* n_records = nullptr;
		unsigned* clip_tiles = nullptr;      // tiles with pixels owned by clipped faces (k_tile_shade -> k_shade_clipped)
		unsigned* n_clip_tiles = nullptr;
		unsigned* clip_faces = nullptr;      // faces that need the clipper (k_setup_raster -> k_setup_clipped); one entry per face of the largest mesh drawn
		unsigned long long clip_cap = 0;
		unsigned* n_clip_faces = nullptr;
		DrawStatus* d_status = nullptr;
		DrawStatus* h_status = nullptr;      // pinned + mapped: written by k_scan_tiles
		DrawStatus* h_status_dev = nullptr;  // device-side alias of h_status
		cudaEvent_t status_event = nullptr;  // geom_stream: counters published
		cudaEvent_t geom_done = nullptr;     // geom_stream: bins complete, the tile kernel may start
		cudaEvent_t shade_done = nullptr;    // main stream: the tile kernel has consumed (and reset) this slot
		bool used = false;
	} slot[2];
	unsigned draw_counter = 0;
	cudaStream_t geom_stream = nullptr;
	unsigned* dirty_map = nullptr;  // axr_set_dirty_map
	bool fill = false;              // axr_set_output_fill
	unsigned* stale_list = nullptr; // axr_clear_stale_tiles: [2 counters][entries]
	unsigned stale_cap = 0;
	// The depth plane `fresh_target` was cleared to +inf by axr_clear (or at creation) and nothing has been drawn into it since: the
	// first draw's merge test `z < fbZ` passes for every drawable z, so it does not read the plane (13 MB of the C3 frame). Off for
	// good once the raw framebuffer pointers have been handed out (axr_framebuffer_device / _ipc: others may write behind our back).
	bool depth_fresh = false, depth_external = false;
	const float* fresh_target = nullptr;
	bool out_rows = false;          // axr_set_output_rows
	uint32_t fill_color = 0;
	float fill_depth = 0.f;
	bool color_fast = true;  // axr_set_color_math: fused colour arithmetic in the shading stage (default) or the reference's individually rounded one
	bool overlap = false;  // axr_set_overlap: geometry stages on geom_stream (else everything on the main stream)
	PendingDraw pending;
	axr_stats stats{};

	// optional per-kernel timing
	bool profiling = false;
	std::vector<cudaEvent_t> prof_events;  // 2 * AXR_NUM_STAGES per profiled draw: (begin, end) of every stage on its own stream
	std::vector<cudaEvent_t> prof_pool;

	std::vector<DeviceMesh> meshes;
	std::vector<DeviceTexture> textures;
	std::vector<void*> ipc_opened;
	std::vector<void*> shared_allocs;
	std::vector<std::pair<void*, size_t>> registered;  // host ranges page-locked by axr_draw_mesh_host
	const float* depth_read_override = nullptr;        // set for the duration of one axr_draw_mesh_host
	int read_depth = 1;

	// axr_draw_mesh_host: the host depth goes up in row chunks on its own stream and the tile kernel is launched once per chunk,
	// so that the upload of chunk b+1 (PCIe host -> device) runs beside the tile kernel of chunk b, whose zero-copy stores use the
	// other PCIe direction
	static constexpr int MAX_HOST_CHUNKS = 4;
	cudaStream_t up_stream = nullptr;
	cudaEvent_t up_done[MAX_HOST_CHUNKS] = {};
	cudaEvent_t depth_free = nullptr;  // main stream: the previous user of the device depth copy is done
	int host_chunks = 0;               // > 0 only while axr_draw_mesh_host issues its draw
	// axr_draw_mesh_host: no depth upload at all, the merge test reads the host depth of the visible pixels through the zero-copy
	// mapping (128 B row reads over PCIe; C3: 0.976 -> 0.874 ms per call). AXR_B200_HOST_DEPTH_ZEROCOPY=0 at axr_create restores the
	// chunked upload of the whole depth plane.
	bool host_depth_zero_copy = true;
	int chunk_ty[MAX_HOST_CHUNKS + 1] = {};  // GPU tile rows [chunk_ty[b], chunk_ty[b+1]) of chunk b

	// depth peeling (draws with a shader that may discard): per-pixel floor keys + the "another pass" flag; allocated on first use
	unsigned long long* peel_floor = nullptr;
	unsigned* peel_again = nullptr;
};

namespace {

int fail(axr_ctx* ctx, int code, const char* fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	if (ctx) ctx->error = buf; else g_create_error = buf;
	return code;
}

#define CU(call)                                                                                          \
	do {                                                                                                  \
		cudaError_t e_ = (call);                                                                          \
		if (e_ != cudaSuccess) return fail(ctx, AXR_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

inline int grid_for(size_t n, int block, int cap = 148 * 16) {
	size_t g = (n + block - 1) / block;
	if (g < 1) g = 1;
	if (g > (size_t)cap) g = cap;
	return (int)g;
}

int n_tiles(const axr_ctx* c) { return c->fp.ntx * c->fp.nty; }

int reset_raster_state(axr_ctx* ctx, int si) {
	axr_ctx::DrawSlot& sl = ctx->slot[si];
	size_t npx = (size_t)ctx->fp.W * ctx->fp.H;
	k_fill_u64<<<grid_for(npx, 256), 256, 0, ctx->stream>>>(sl.vis, KEY_EMPTY, npx);
	k_fill_u32<<<grid_for(n_tiles(ctx), 256), 256, 0, ctx->stream>>>(sl.tile_touched, 0u, (size_t)n_tiles(ctx));
	k_fill_u32<<<grid_for(n_tiles(ctx), 256), 256, 0, ctx->stream>>>(sl.tile_count, 0u, (size_t)n_tiles(ctx));
	CU(cudaGetLastError());
	return AXR_OK;
}

int ensure_bins(axr_ctx* ctx, int si, unsigned want_rec, unsigned want_ref) {
	axr_ctx::DrawSlot& sl = ctx->slot[si];
	if (want_rec > sl.rec_cap) {
		if (sl.records) CU(cudaFree(sl.records));
		sl.records = nullptr;
		unsigned cap = want_rec + want_rec / 8 + 1024;
		CU(cudaMalloc(&sl.records, (size_t)cap * sizeof(TriRecord)));
		sl.rec_cap = cap;
	}
	if (want_ref > sl.ref_cap) {
		if (sl.items) CU(cudaFree(sl.items));
		sl.items = nullptr;
		unsigned cap = want_ref + want_ref / 8 + 4096;
		CU(cudaMalloc(&sl.items, (size_t)cap * sizeof(unsigned)));
		sl.ref_cap = cap;
	}
	return AXR_OK;
}

int sync_materials(axr_ctx* ctx, DeviceMesh& m) {
	if (!m.materials_dirty) return AXR_OK;
	CU(cudaStreamSynchronize(ctx->geom_stream));  // no draw in flight may still read the old table
	CU(cudaStreamSynchronize(ctx->stream));
	CU(cudaMemcpyAsync(m.d_materials, m.materials.data(), m.materials.size() * sizeof(Material), cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));  // the host vector may change right after
	m.materials_dirty = false;
	return AXR_OK;
}

int issue_draw(axr_ctx* ctx, axr_mesh mh, const float* model, int si, bool peel);

// Inspect the status of the draw issued last; if its bins overflowed (its tile kernel then did nothing), grow and re-issue.
// Every API call that touches the context starts here, so at most one draw is ever unchecked and a later draw is never
// composited ahead of an earlier one that has to be redone.
int check_pending(axr_ctx* ctx) {
	if (!ctx->pending.valid) return AXR_OK;
	PendingDraw p = ctx->pending;
	axr_ctx::DrawSlot& sl = ctx->slot[p.slot];
	CU(cudaEventSynchronize(sl.status_event));
	ctx->pending.valid = false;
	struct { unsigned long long clipped_faces, triangles, small_triangles, binned_triangles, bin_refs; unsigned overflow, pad; } st;
	static_assert(sizeof(st) == 48 && offsetof(DrawStatus, stripes) == 48, "status head layout");
	memcpy(&st, sl.h_status, sizeof st);
	ctx->stats.clipped_faces = st.clipped_faces;
	ctx->stats.triangles = st.triangles;
	ctx->stats.small_triangles = st.small_triangles;
	ctx->stats.binned_triangles = st.binned_triangles;
	ctx->stats.bin_refs = st.bin_refs;
	DeviceMesh& pm = ctx->meshes[p.mesh];
	if (!st.overflow) {
		// learn whether this mesh needs the bin kernels at all (BASELINE configs 2-4: every triangle is rasterised by its setup thread)
		if (!p.peel) pm.bin_mode = st.binned_triangles == 0 ? BINS_NONE : (st.binned_triangles <= BINS_SCAN_MAX / 2 ? BINS_SCAN : BINS_LISTS);
		return AXR_OK;
	}
	if (st.overflow & OVF_NEED_BINS) pm.bin_mode = BINS_LISTS;
	if (p.redo_depth >= 3) return fail(ctx, AXR_ERR_CAPACITY, "bin capacity still exceeded after regrowing (records %llu refs %llu)",
	                                   (unsigned long long)st.binned_triangles, (unsigned long long)st.bin_refs);
	CU(cudaStreamSynchronize(ctx->geom_stream));
	CU(cudaStreamSynchronize(ctx->stream));
	int rc = ensure_bins(ctx, p.slot, (unsigned)st.binned_triangles, (unsigned)st.bin_refs);
	if (rc) return rc;
	rc = reset_raster_state(ctx, p.slot);
	if (rc) return rc;
	CU(cudaStreamSynchronize(ctx->stream));
	rc = issue_draw(ctx, p.mesh, p.model, p.slot, p.peel);
	if (rc) return rc;
	ctx->pending.redo_depth = p.redo_depth + 1;
	ctx->stats.redo = 1;
	return check_pending(ctx);
}

cudaEvent_t prof_mark(axr_ctx* ctx, cudaStream_t s) {
	if (!ctx->profiling) return nullptr;
	cudaEvent_t e = nullptr;
	if (!ctx->prof_pool.empty()) { e = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
	else if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
	cudaEventRecord(e, s);
	ctx->prof_events.push_back(e);
	return e;
}


// which: 0 = k_tile_shade over a gx x gy grid of tiles, 1 = k_shade_clipped (fixed grid)
using TileLaunch = int (*)(axr_ctx*, const MeshView&, const Uniforms&, const FrameParams&, const TileIn&, int which, unsigned gx, unsigned gy);
template <typename Shader, int SMP, bool FAST>
int launch_builtin(axr_ctx* ctx, const MeshView& mv, const Uniforms& u, const FrameParams& fp, const TileIn& in, int which, unsigned gx, unsigned gy) {
	if (which == 0) {
		if constexpr (!Shader::DISCARDS) {  // axr_set_output_fill is refused for shaders that discard
			if (in.fill) {
				k_tile_shade<Shader, SMP, FAST, true><<<dim3(gx, gy), TILE_THREADS, 0, ctx->stream>>>(mv, u, fp, in);
				return AXR_OK;
			}
		}
		k_tile_shade<Shader, SMP, FAST, false><<<dim3(gx, gy), TILE_THREADS, 0, ctx->stream>>>(mv, u, fp, in);
	} else {
		k_shade_clipped<Shader, SMP><<<CLIP_SHADE_CTAS, CLIP_SHADE_THREADS, 0, ctx->stream>>>(mv, u, fp, in);
	}
	return AXR_OK;
}
int launch_plugin(axr_ctx* ctx, const MeshView& mv, const Uniforms& u, const FrameParams& fp, const TileIn& in, int which, unsigned gx, unsigned gy) {
	const ShaderPlugin& p = ctx->plugins[ctx->shader_kind - AXR_SHADER_PLUGIN_BASE];
	const int e = p.launch(&mv, &u, &fp, &in, ctx->device, u.sampler, which, gx, gy, ctx->stream);
	if (e) return fail(ctx, AXR_ERR_CUDA, "shader plug-in %s: launch failed: %s", p.path.c_str(), cudaGetErrorString((cudaError_t)e));
	return AXR_OK;
}
template <typename Shader>
TileLaunch builtin_launcher(const axr_ctx* ctx, int sampler) {
	const bool fast = ctx->color_fast && Shader::HAS_FAST;
	if (sampler) return fast ? &launch_builtin<Shader, 1, true> : &launch_builtin<Shader, 1, false>;
	return fast ? &launch_builtin<Shader, 0, true> : &launch_builtin<Shader, 0, false>;
}

// One launch over all tile rows of the band, or (axr_draw_mesh_host) one launch per uploaded row chunk, each behind its upload;
// then the small kernel for the pixels owned by clipped faces (k_shade_clipped: a fixed grid over a list that is usually empty).
int launch_tile_kernels(axr_ctx* ctx, TileLaunch fn, const MeshView& mv, const Uniforms& u, const TileIn& in, uint64_t& launches) {
	const int chunks = ctx->host_chunks > 0 ? ctx->host_chunks : 1;
	for (int b = 0; b < chunks; ++b) {
		FrameParams fp = ctx->fp;
		if (ctx->host_chunks > 0) {
			fp.ty_lo = ctx->chunk_ty[b]; fp.ty_hi = ctx->chunk_ty[b + 1];
			cudaStreamWaitEvent(ctx->stream, ctx->up_done[b], 0);
		}
		if (int rc = fn(ctx, mv, u, fp, in, 0, (unsigned)fp.ntx, (unsigned)(fp.ty_hi - fp.ty_lo))) return rc;
	}
	if (int rc = fn(ctx, mv, u, ctx->fp, in, 1, 0, 0)) return rc;
	launches += chunks + 1;
	return AXR_OK;
}

bool shader_is_plugin(const axr_ctx* ctx, int kind) { return kind >= AXR_SHADER_PLUGIN_BASE && (size_t)(kind - AXR_SHADER_PLUGIN_BASE) < ctx->plugins.size(); }
// fragment() may return true: the draw is depth-peeled
bool shader_discards(const axr_ctx* ctx) {
	if (ctx->shader_kind == AXR_SHADER_CUTOUT) return true;
	return shader_is_plugin(ctx, ctx->shader_kind) && ctx->plugins[ctx->shader_kind - AXR_SHADER_PLUGIN_BASE].discards;
}

// One pass of the five kernels. peel: the pass belongs to a depth-peeled draw (draw_peeled below) — the raster sites reject
// keys at or below the pixel's floor, and everything stays on the main stream (the floor buffer is ordered on it).
int issue_draw(axr_ctx* ctx, axr_mesh mh, const float* model, int si, bool peel) {
	DeviceMesh& m = ctx->meshes[mh];
	axr_ctx::DrawSlot& sl = ctx->slot[si];
	if (peel != shader_discards(ctx)) return fail(ctx, AXR_ERR_INVALID, "internal: peel flag does not match the shader");
	if (ctx->fill && (peel || ctx->host_chunks > 0))
		return fail(ctx, AXR_ERR_UNSUPPORTED, "axr_set_output_fill: not for shaders that discard (several passes per draw) nor for host framebuffers");
	int rc = sync_materials(ctx, m);
	if (rc) return rc;
	// shader / material validation (the reference dereferences null textures, include/shaders/shaders.hpp:178,210)
	for (const Material& mat : m.materials) {
		if ((ctx->shader_kind == AXR_SHADER_PHONG || ctx->shader_kind == AXR_SHADER_PBR) && (!mat.tex[0].data || !mat.tex[1].data))
			return fail(ctx, AXR_ERR_MATERIAL, "shader needs diffuse + bump textures on every material group");
		if (ctx->shader_kind == AXR_SHADER_CUTOUT && !mat.tex[0].data)
			return fail(ctx, AXR_ERR_MATERIAL, "CutoutShader needs a diffuse texture on every material group");
		if (ctx->shader_kind == AXR_SHADER_PBR && (!mat.tex[2].data || !mat.tex[3].data || !mat.tex[4].data))
			return fail(ctx, AXR_ERR_MATERIAL, "PBRShader needs metallic, roughness and ao textures on every material group");
		if (shader_is_plugin(ctx, ctx->shader_kind))
			for (int k = 0; k < 5; ++k)
				if (((ctx->plugins[ctx->shader_kind - AXR_SHADER_PLUGIN_BASE].textures >> k) & 1u) && !mat.tex[k].data)
					return fail(ctx, AXR_ERR_MATERIAL, "the plug-in shader needs texture slot %d on every material group", k);
	}
	Uniforms u;
	u.model = load_m4(model);
	u.mvp = mul(load_m4(ctx->view_proj), u.model);  // reference src/tiled_pipeline.cpp:149
	m4 inv = inverse(u.model);
	u.normal_mat.c[0] = V3(inv.c[0].x, inv.c[1].x, inv.c[2].x);  // mat3(transpose(inverse(model)))
	u.normal_mat.c[1] = V3(inv.c[0].y, inv.c[1].y, inv.c[2].y);
	u.normal_mat.c[2] = V3(inv.c[0].z, inv.c[1].z, inv.c[2].z);
	u.cam_pos = V3(ctx->cam_pos[0], ctx->cam_pos[1], ctx->cam_pos[2]);
	u.light_dir = V3(ctx->shader_params.light_dir[0], ctx->shader_params.light_dir[1], ctx->shader_params.light_dir[2]);
	u.light_color = V3(ctx->shader_params.light_color[0], ctx->shader_params.light_color[1], ctx->shader_params.light_color[2]);
	u.sampler = ctx->sampler;
	memcpy(u.user, ctx->shader_user, sizeof u.user);

	MeshView mv;
	mv.pos = m.pos; mv.attr = m.attr; mv.idx = m.idx; mv.idx4 = m.idx4;
	mv.n_plane = m.n_verts ? m.n_verts : 1;
	mv.n_verts = m.n_verts; mv.n_faces = m.n_faces;
	mv.materials = m.d_materials;
	mv.material0 = m.materials.empty() ? Material{} : m.materials[0];
	mv.group_first = m.d_group_first;
	mv.n_groups = (int)m.materials.size();
	mv.first_face = (unsigned)m.group_first[0];

	// ---- geometry stages on geom_stream. The slot was last used two draws ago: wait until that draw's tile kernel has
	//      consumed it (it resets the keys, flags and cursors it read).
	cudaStream_t g = (ctx->overlap && !peel) ? ctx->geom_stream : ctx->stream;
	if (sl.clip_cap < m.n_faces) {  // first draw of a mesh larger than any before it on this slot
		CU(cudaStreamSynchronize(ctx->geom_stream));
		CU(cudaStreamSynchronize(ctx->stream));
		if (sl.clip_faces) CU(cudaFree(sl.clip_faces));
		sl.clip_faces = nullptr; sl.clip_cap = 0;
		CU(cudaMalloc(&sl.clip_faces, (size_t)m.n_faces * 4));
		sl.clip_cap = m.n_faces;
	}
	if (sl.used) CU(cudaStreamWaitEvent(g, sl.shade_done, 0));
	uint64_t launches = 0;
	prof_mark(ctx, g);
	{
		// at least sizeof(DrawStatus)/4 threads: the kernel also zeroes the draw's counters
		const unsigned long long per = (m.n_verts + AXR_VERTEX_PER_THREAD - 1) / AXR_VERTEX_PER_THREAD;
		const unsigned long long threads = per > sizeof(DrawStatus) / 4 ? per : sizeof(DrawStatus) / 4;
		k_vertex_xform<<<(unsigned)((threads + 255) / 256), 256, 0, g>>>(m.pos, m.n_verts, u.mvp, (float)ctx->fp.W, (float)ctx->fp.H, m.sv[si],
		                                                                sl.d_status, sl.n_records, sl.n_clip_tiles, sl.n_clip_faces);
		++launches;
	}
	prof_mark(ctx, g);
	prof_mark(ctx, g);
	SetupOut so;
	so.vis = sl.vis; so.tile_touched = sl.tile_touched; so.tile_count = sl.tile_count;
	so.records = sl.records; so.rec_cap = sl.rec_cap; so.n_records = sl.n_records; so.status = sl.d_status;
	const bool tput = m.n_faces >= SMALL_TPUT_MIN_FACES;
	so.small_dim = tput ? SMALL_DIM_TPUT : SMALL_DIM_LAT;
	so.small_area = tput ? SMALL_AREA_TPUT : SMALL_AREA_LAT;
	so.floor = peel ? ctx->peel_floor : nullptr;
	so.clip_faces = sl.clip_faces; so.n_clip_faces = sl.n_clip_faces;
	so.banded = (ctx->fp.y_lo > 0 || ctx->fp.y_hi < ctx->fp.H) ? 1 : 0;
	so.band_lo = (float)ctx->fp.y_lo; so.band_hi = (float)ctx->fp.y_hi;
	so.n_chunks = 0; so.swz_rows = 0;
	const int bin_mode = peel ? (int)BINS_LISTS : m.bin_mode;
	const bool bins = bin_mode == BINS_LISTS;
	so.bins_enabled = bin_mode;
	{
		const unsigned grid = m.n_faces ? setup_grid(m.n_faces, so.n_chunks, so.swz_rows) : 0u;
		if (grid) {
			if (peel) k_setup_raster<true><<<grid, SETUP_THREADS, 0, g>>>(mv, m.sv[si], ctx->fp, so);
			else k_setup_raster<false><<<grid, SETUP_THREADS, 0, g>>>(mv, m.sv[si], ctx->fp, so);
			++launches;
		}
		// the faces that need the clipper + (its last CTA) the fold of the draw's counters into the status the host reads
		if (peel) k_setup_clipped<true><<<CLIPSETUP_CTAS, CLIPSETUP_THREADS, 0, g>>>(mv, u.mvp, ctx->fp, so, sl.h_status_dev);
		else k_setup_clipped<false><<<CLIPSETUP_CTAS, CLIPSETUP_THREADS, 0, g>>>(mv, u.mvp, ctx->fp, so, sl.h_status_dev);
		++launches;
	}
	prof_mark(ctx, g);
	prof_mark(ctx, g);
	if (bins) {
		k_scan_tiles<<<1, SCAN_THREADS, 0, g>>>(sl.tile_count, sl.bin_start, n_tiles(ctx), sl.ref_cap, sl.n_records, sl.rec_cap, sl.d_status,
		                                       sl.h_status_dev);
		++launches;
	}
	prof_mark(ctx, g);
	CU(cudaEventRecord(sl.status_event, g));  // the status has been stored into mapped host memory (by the setup kernel's last CTA / the scan)
	prof_mark(ctx, g);
	if (bins) {
		k_bin_scatter<<<148 * 4, 256, 0, g>>>(sl.records, sl.n_records, ctx->fp, sl.bin_start, sl.tile_count, sl.items, sl.d_status);
		++launches;
	}
	prof_mark(ctx, g);
	CU(cudaEventRecord(sl.geom_done, g));
	// ---- tile raster + shading + resolve on the main stream (the one clears, uploads and resolves are ordered on)
	CU(cudaStreamWaitEvent(ctx->stream, sl.geom_done, 0));
	prof_mark(ctx, ctx->stream);
	TileIn in;
	in.vis = sl.vis; in.tile_touched = sl.tile_touched; in.tile_cursor = sl.tile_count; in.bin_start = sl.bin_start;
	in.items = sl.items; in.records = sl.records; in.n_records = sl.n_records; in.status = sl.d_status; in.sv = m.sv[si];
	in.color = ctx->out_color; in.depth = ctx->out_depth; in.read_depth = ctx->read_depth;
	if (ctx->depth_fresh && !ctx->depth_external && ctx->fresh_target == ctx->out_depth && ctx->out_depth == ctx->depth && ctx->host_chunks == 0 && !peel)
		in.read_depth = 0;  // (own plane only: a redirected output may have other writers)
	ctx->depth_fresh = false;  // whatever this draw is, the plane is no longer known to be all +inf
	in.depth_read = ctx->depth_read_override ? ctx->depth_read_override : ctx->out_depth;
	in.row_major = (ctx->host_chunks > 0 || ctx->out_rows) ? 1 : 0;
	in.floor = peel ? ctx->peel_floor : nullptr;
	in.again = peel ? ctx->peel_again : nullptr;
	in.clip_tiles = sl.clip_tiles; in.n_clip_tiles = sl.n_clip_tiles;
	in.dirty = ctx->dirty_map;
	in.bin_mode = bin_mode;
	in.fill = ctx->fill ? 1 : 0; in.fill_color = ctx->fill_color; in.fill_depth = ctx->fill_depth;
	TileLaunch fn = nullptr;
	switch (ctx->shader_kind) {
	case AXR_SHADER_FLAT: fn = builtin_launcher<FlatShader>(ctx, u.sampler); break;
	case AXR_SHADER_PHONG: fn = builtin_launcher<PhongShader>(ctx, u.sampler); break;
	case AXR_SHADER_PBR: fn = builtin_launcher<PBRShader>(ctx, u.sampler); break;
	case AXR_SHADER_CUTOUT: fn = builtin_launcher<CutoutShader>(ctx, u.sampler); break;
	default:
		if (!shader_is_plugin(ctx, ctx->shader_kind)) return fail(ctx, AXR_ERR_UNSUPPORTED, "unknown shader kind %d", ctx->shader_kind);
		fn = &launch_plugin;
	}
	rc = launch_tile_kernels(ctx, fn, mv, u, in, launches);
	if (rc) return rc;
	prof_mark(ctx, ctx->stream);
	CU(cudaEventRecord(sl.shade_done, ctx->stream));
	sl.used = true;
	CU(cudaGetLastError());
	ctx->stats.faces = m.n_faces;
	ctx->stats.kernel_launches = launches;
	ctx->stats.redo = 0;
	ctx->pending.valid = true;
	ctx->pending.slot = si;
	ctx->pending.mesh = mh;
	memcpy(ctx->pending.model, model, sizeof(float) * 16);
	ctx->pending.redo_depth = 0;
	ctx->pending.peel = peel;
	ctx->pending.bins = bins;
	return AXR_OK;
}

// Draw with a shader whose fragment() may discard (reference src/tiled_pipeline.cpp:569-577: the fragment shader runs for
// every fragment that passes the early-Z test, in triangle order, and a discarded one leaves depth and colour alone). The
// pixel's final owner is therefore the smallest (z, order) key among its NON-discarded fragments, which the deferred
// visibility pass cannot know in advance. Depth peeling finds it without ordering anything: pass n keeps, per pixel, the
// smallest key above that pixel's floor; the tile kernel shades it; if the shader discards it the key becomes the new
// floor and another pass is needed, otherwise the pixel is finished (floor = KEY_EMPTY, which no key exceeds). Passes
// repeat until no winner was discarded: 1 + (deepest run of discarded fragments in front of a pixel's owner) passes.
constexpr int MAX_PEEL_PASSES = 256;
int draw_peeled(axr_ctx* ctx, axr_mesh mh, const float* model, int si) {
	const size_t npx = (size_t)ctx->fp.W * ctx->fp.H;
	if (!ctx->peel_floor) {
		CU(cudaMalloc(&ctx->peel_floor, npx * 8));
		CU(cudaMalloc(&ctx->peel_again, 4));
	}
	k_fill_u64<<<grid_for(npx, 256), 256, 0, ctx->stream>>>(ctx->peel_floor, 0ull, npx);  // every key is above 0
	uint64_t launches = 1;
	for (int pass = 0; pass < MAX_PEEL_PASSES; ++pass) {
		CU(cudaMemsetAsync(ctx->peel_again, 0, 4, ctx->stream));
		int rc = issue_draw(ctx, mh, model, si, true);
		if (rc) return rc;
		rc = check_pending(ctx);  // bins overflowed: regrown and this pass re-issued (its tile kernel had done nothing)
		if (rc) return rc;
		launches += ctx->stats.kernel_launches;
		unsigned again = 0;
		CU(cudaMemcpyAsync(&again, ctx->peel_again, 4, cudaMemcpyDeviceToHost, ctx->stream));
		CU(cudaStreamSynchronize(ctx->stream));
		if (!again) {
			ctx->stats.kernel_launches = launches;
			return AXR_OK;
		}
	}
	return fail(ctx, AXR_ERR_CAPACITY, "more than %d discarded fragments in front of one pixel", MAX_PEEL_PASSES - 1);
}

// Entry point of every draw call: picks the raster slot and the plain or the peeled path.
int draw(axr_ctx* ctx, axr_mesh mh, const float* model) {
	const int si = (int)(ctx->draw_counter++ & 1u);
	if (shader_discards(ctx)) return draw_peeled(ctx, mh, model, si);
	return issue_draw(ctx, mh, model, si, false);
}

// Every host-visible synchronisation waits for both streams.
int sync_all(axr_ctx* ctx) {
	CU(cudaStreamSynchronize(ctx->geom_stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return AXR_OK;
}

bool valid_mesh(const axr_ctx* ctx, axr_mesh m) { return m >= 0 && (size_t)m < ctx->meshes.size() && ctx->meshes[m].live; }
bool valid_tex(const axr_ctx* ctx, axr_tex t) { return t >= 0 && (size_t)t < ctx->textures.size() && ctx->textures[t].live; }

}  // namespace

extern "C" {

int axr_abi_version(void) { return AXR_ABI_VERSION; }

const char* axr_last_error(const axr_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int axr_create(const axr_config* cfg, axr_ctx** out) {
	axr_ctx* ctx = nullptr;  // for the CU/fail macros: errors land in g_create_error
	if (!cfg || !out) return fail(ctx, AXR_ERR_INVALID, "axr_create: null argument");
	*out = nullptr;
	if (cfg->width <= 0 || cfg->height <= 0 || cfg->width > 65536 || cfg->height > 65536)
		return fail(ctx, AXR_ERR_INVALID, "axr_create: bad framebuffer size %dx%d", cfg->width, cfg->height);
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
		return fail(ctx, AXR_ERR_NO_DEVICE, "axr_create: no CUDA device (this library has no CPU path)");
	if (cfg->device < 0 || cfg->device >= ndev) return fail(ctx, AXR_ERR_INVALID, "axr_create: device %d of %d", cfg->device, ndev);
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, cfg->device));
	if (prop.major != 10) return fail(ctx, AXR_ERR_NO_DEVICE, "axr_create: device %d is sm_%d%d, kernels are built for sm_100a only", cfg->device, prop.major, prop.minor);
	int y0 = cfg->band_y0, y1 = cfg->band_y1;
	if (y0 == 0 && y1 == 0) y1 = cfg->height;
	if (y0 < 0 || y1 > cfg->height || y0 >= y1 || (y0 % REF_TILE) != 0)
		return fail(ctx, AXR_ERR_INVALID, "axr_create: bad band [%d,%d) (start must be a multiple of %d)", y0, y1, REF_TILE);
	CU(cudaSetDevice(cfg->device));
	axr_ctx* c = new axr_ctx();
	ctx = c;
	c->device = cfg->device;
	c->fp.W = cfg->width; c->fp.H = cfg->height;
	c->fp.y_lo = y0; c->fp.y_hi = y1;
	c->fp.ntx = (cfg->width + GT - 1) / GT;
	c->fp.nty = (cfg->height + GT - 1) / GT;
	c->fp.ty_lo = y0 / GT;
	c->fp.ty_hi = (y1 + GT - 1) / GT;
	c->sampler = cfg->sampler ? 1 : 0;
	if (const char* e = getenv("AXR_B200_HOST_DEPTH_ZEROCOPY")) c->host_depth_zero_copy = e[0] == '1';
	if (const char* e = getenv("AXR_B200_COLOR_MATH")) c->color_fast = strcmp(e, "exact") != 0;
	// identity uniforms until axr_set_uniforms
	memset(c->view_proj, 0, sizeof c->view_proj);
	memset(c->viewport, 0, sizeof c->viewport);
	for (int i = 0; i < 4; ++i) c->view_proj[i * 5] = c->viewport[i * 5] = 1.0f;
	c->cam_pos[0] = c->cam_pos[1] = c->cam_pos[2] = 0.f;
	c->shader_params.light_dir[1] = -1.0f;
	c->shader_params.light_color[0] = c->shader_params.light_color[1] = c->shader_params.light_color[2] = 1.0f;
#define CUC(call)                                                                              \
	do {                                                                                       \
		cudaError_t e_ = (call);                                                               \
		if (e_ != cudaSuccess) {                                                               \
			int rc_ = fail(nullptr, AXR_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));    \
			axr_destroy(c);                                                                    \
			return rc_;                                                                        \
		}                                                                                      \
	} while (0)
	if (cfg->stream) { c->stream = (cudaStream_t)cfg->stream; c->own_stream = false; }
	else { CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
	const size_t npx = (size_t)cfg->width * cfg->height;
	const size_t nt = (size_t)c->fp.ntx * c->fp.nty;
	CUC(cudaMalloc(&c->color, npx * 4));
	CUC(cudaMalloc(&c->depth, npx * 4));
	c->out_color = c->color; c->out_depth = c->depth;
	{
		// geometry CTAs are scheduled ahead of the (much longer) tile kernel's remaining CTAs, so the two really interleave on the SMs
		int lo = 0, hi = 0;
		cudaDeviceGetStreamPriorityRange(&lo, &hi);
		(void)lo;
		CUC(cudaStreamCreateWithPriority(&c->geom_stream, cudaStreamNonBlocking, hi));
	}
	CUC(cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
	for (auto& e : c->up_done) CUC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	CUC(cudaEventCreateWithFlags(&c->depth_free, cudaEventDisableTiming));
	for (int si = 0; si < 2; ++si) {
		axr_ctx::DrawSlot& sl = c->slot[si];
		CUC(cudaMalloc(&sl.vis, npx * 8));
		CUC(cudaMalloc(&sl.tile_touched, nt * 4));
		CUC(cudaMalloc(&sl.tile_count, nt * 4));
		CUC(cudaMalloc(&sl.bin_start, (nt + 1) * 4));
		CUC(cudaMalloc(&sl.n_records, 4));
		CUC(cudaMalloc(&sl.clip_tiles, nt * 4));
		CUC(cudaMalloc(&sl.n_clip_tiles, 4));
		CUC(cudaMemsetAsync(sl.n_clip_tiles, 0, 4, c->stream));
		CUC(cudaMalloc(&sl.n_clip_faces, 4));
		CUC(cudaMemsetAsync(sl.n_clip_faces, 0, 4, c->stream));
		CUC(cudaMalloc(&sl.d_status, sizeof(DrawStatus)));
		CUC(cudaHostAlloc(&sl.h_status, 64, cudaHostAllocMapped));
		CUC(cudaHostGetDevicePointer(&sl.h_status_dev, sl.h_status, 0));
		CUC(cudaEventCreateWithFlags(&sl.status_event, cudaEventDisableTiming));
		CUC(cudaEventCreateWithFlags(&sl.geom_done, cudaEventDisableTiming));
		CUC(cudaEventCreateWithFlags(&sl.shade_done, cudaEventDisableTiming));
		sl.rec_cap = 1u << 18; sl.ref_cap = 1u << 20;
		CUC(cudaMalloc(&sl.records, (size_t)sl.rec_cap * sizeof(TriRecord)));
		CUC(cudaMalloc(&sl.items, (size_t)sl.ref_cap * 4));
	}
	for (int si = 0; si < 2; ++si)
		if (reset_raster_state(c, si) != AXR_OK) { g_create_error = c->error; axr_destroy(c); return AXR_ERR_CUDA; }
	k_clear<<<grid_for(npx, 256), 256, 0, c->stream>>>(c->color, c->depth, 0u, INFINITY, 0, npx);  // Framebuffer ctor: colour 0, depth +inf
	c->depth_fresh = true; c->fresh_target = c->depth;
	CUC(cudaStreamSynchronize(c->stream));
#undef CUC
	*out = c;
	return AXR_OK;
}

void axr_destroy(axr_ctx* ctx) {
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	if (ctx->geom_stream) cudaStreamSynchronize(ctx->geom_stream);
	if (ctx->stream) cudaStreamSynchronize(ctx->stream);
	for (auto& m : ctx->meshes) if (m.live) { cudaFree(m.pos); cudaFree(m.attr); cudaFree(m.idx); cudaFree(m.idx4); cudaFree(m.sv[0]); cudaFree(m.sv[1]); cudaFree(m.d_materials); cudaFree(m.d_group_first); }
	// shader plug-ins stay mapped: unloading a library with its own static CUDA runtime while the process lives on is not safe
	for (auto& t : ctx->textures) if (t.live) cudaFree(t.data);
	for (void* p : ctx->ipc_opened) cudaIpcCloseMemHandle(p);
	for (void* p : ctx->shared_allocs) cudaFree(p);
	for (auto& r : ctx->registered) cudaHostUnregister(r.first);
	cudaFree(ctx->color); cudaFree(ctx->depth);
	cudaFree(ctx->peel_floor); cudaFree(ctx->peel_again);
	cudaFree(ctx->stale_list);
	for (auto& sl : ctx->slot) {
		cudaFree(sl.vis); cudaFree(sl.tile_touched); cudaFree(sl.tile_count); cudaFree(sl.bin_start); cudaFree(sl.items);
		cudaFree(sl.records); cudaFree(sl.n_records); cudaFree(sl.d_status); cudaFree(sl.clip_tiles); cudaFree(sl.n_clip_tiles);
		cudaFree(sl.clip_faces); cudaFree(sl.n_clip_faces);
		if (sl.h_status) cudaFreeHost(sl.h_status);
		if (sl.status_event) cudaEventDestroy(sl.status_event);
		if (sl.geom_done) cudaEventDestroy(sl.geom_done);
		if (sl.shade_done) cudaEventDestroy(sl.shade_done);
	}
	if (ctx->geom_stream) cudaStreamDestroy(ctx->geom_stream);
	if (ctx->up_stream) { cudaStreamSynchronize(ctx->up_stream); cudaStreamDestroy(ctx->up_stream); }
	for (cudaEvent_t e : ctx->up_done) if (e) cudaEventDestroy(e);
	if (ctx->depth_free) cudaEventDestroy(ctx->depth_free);
	for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
	for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
	if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
	delete ctx;
}

int axr_upload_mesh(axr_ctx* ctx, const float* vertices, uint64_t n_verts, const uint32_t* indices, uint64_t n_faces,
                    const axr_group* groups, uint32_t n_groups, axr_mesh* out) {
	if (!ctx) return AXR_ERR_INVALID;
	if (!out || (n_verts && !vertices) || (n_faces && !indices)) return fail(ctx, AXR_ERR_INVALID, "axr_upload_mesh: null argument");
	if (n_faces >= (1ull << 29)) return fail(ctx, AXR_ERR_CAPACITY, "axr_upload_mesh: %llu faces exceed the 2^29 ordinal range", (unsigned long long)n_faces);
	if (n_verts >= (1ull << 32)) return fail(ctx, AXR_ERR_CAPACITY, "axr_upload_mesh: %llu vertices exceed 32-bit indices", (unsigned long long)n_verts);
	for (uint64_t i = 0; i < n_faces * 3; ++i)
		if (indices[i] >= n_verts) return fail(ctx, AXR_ERR_INVALID, "axr_upload_mesh: index %u out of range at face %llu", indices[i], (unsigned long long)(i / 3));
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	DeviceMesh m;
	m.n_verts = n_verts; m.n_faces = n_faces;
	// material groups: contiguous ascending face ranges (reference src/mesh.cpp:336-346)
	if (!groups || n_groups == 0) {
		m.group_first = {0ull, (unsigned long long)n_faces};
	} else {
		// (the faces of an OBJ in front of its first `usemtl` belong to no group: drawMesh walks the groups, :176-179, so they are never drawn)
		unsigned long long expect = groups[0].first_face;
		for (uint32_t g = 0; g < n_groups; ++g) {
			if (groups[g].first_face != expect) return fail(ctx, AXR_ERR_INVALID, "axr_upload_mesh: group %u does not start where group %u ends", g, g ? g - 1 : 0);
			m.group_first.push_back(groups[g].first_face);
			expect += groups[g].face_count;
		}
		if (expect != n_faces) return fail(ctx, AXR_ERR_INVALID, "axr_upload_mesh: groups cover %llu of %llu faces", expect, (unsigned long long)n_faces);
		m.group_first.push_back(expect);
	}
	const size_t ng = m.group_first.size() - 1;
	m.materials.assign(ng, Material{});
	const size_t nv = n_verts ? n_verts : 1, nf = n_faces ? n_faces : 1;
	float* raw = nullptr;
	auto release = [&]() {  // a failed upload leaves nothing behind
		cudaFree(m.pos); cudaFree(m.attr); cudaFree(m.sv[0]); cudaFree(m.sv[1]); cudaFree(m.idx); cudaFree(m.idx4); cudaFree(m.d_materials);
		cudaFree(m.d_group_first); cudaFree(raw);
	};
#define CUM(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { release(); return fail(ctx, AXR_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } } while (0)
	CUM(cudaMalloc(&m.pos, nv * sizeof(float4)));
	CUM(cudaMalloc(&m.attr, nv * 3 * sizeof(float4)));
	CUM(cudaMalloc(&m.sv[0], nv * sizeof(float4)));
	CUM(cudaMalloc(&m.sv[1], nv * sizeof(float4)));
	CUM(cudaMalloc(&m.idx, nf * 3 * sizeof(unsigned)));
	CUM(cudaMalloc(&m.d_materials, ng * sizeof(Material)));
	CUM(cudaMalloc(&m.d_group_first, (ng + 1) * sizeof(unsigned long long)));
	if (n_verts) {
		CUM(cudaMalloc(&raw, n_verts * 14 * sizeof(float)));
		CUM(cudaMemcpyAsync(raw, vertices, n_verts * 14 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
		k_split_vertices<<<(unsigned)((n_verts + 255) / 256), 256, 0, ctx->stream>>>(raw, n_verts, nv, m.pos, m.attr);
	}
	if (n_faces) CUM(cudaMemcpyAsync(m.idx, indices, n_faces * 3 * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
#if AXR_IDX_PAD
	CUM(cudaMalloc(&m.idx4, nf * sizeof(uint4)));
	if (n_faces) k_pad_indices<<<(unsigned)((n_faces + 255) / 256), 256, 0, ctx->stream>>>(m.idx, n_faces, m.idx4);
#endif
	CUM(cudaMemcpyAsync(m.d_group_first, m.group_first.data(), (ng + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
	CUM(cudaStreamSynchronize(ctx->stream));
#undef CUM
	if (raw) { cudaFree(raw); raw = nullptr; }
	m.live = true;
	m.materials_dirty = true;
	size_t slot = ctx->meshes.size();
	for (size_t i = 0; i < ctx->meshes.size(); ++i) if (!ctx->meshes[i].live) { slot = i; break; }
	if (slot == ctx->meshes.size()) ctx->meshes.push_back(std::move(m)); else ctx->meshes[slot] = std::move(m);
	*out = (axr_mesh)slot;
	return AXR_OK;
}

// ---- OBJ / MTL ingestion (reference src/mesh.cpp): text parse + de-duplication on the host (axr_obj.hpp), tangents on the device
int axr_load_obj(axr_ctx* ctx, const char* text, size_t len, axr_mesh* out, axr_obj_info* info) {
	if (!ctx) return AXR_ERR_INVALID;
	if (!out || (len && !text)) return fail(ctx, AXR_ERR_INVALID, "axr_load_obj: null argument");
	axr_obj::Parsed ps;
	if (!axr_obj::parse_obj(text, len, ps)) return fail(ctx, AXR_ERR_INVALID, "axr_load_obj: %s", ps.error.c_str());
	const uint64_t nv = ps.v8.size() / 8, nf = ps.idx.size() / 3;
	std::vector<float> v14(nv * 14);
	int rc = axr_generate_tangents(ctx, ps.v8.data(), nv, ps.idx.data(), nf, v14.data());
	if (rc) return rc;
	// groups as the reference's loader leaves them; an OBJ without `usemtl` has none and the reference draws nothing of it
	std::vector<axr_group> groups;
	for (const auto& g : ps.groups) groups.push_back({g.first_face, g.face_count});
	if (groups.empty()) groups.push_back({nf, 0});
	rc = axr_upload_mesh(ctx, v14.data(), nv, ps.idx.data(), nf, groups.data(), (uint32_t)groups.size(), out);
	if (rc) return rc;
	DeviceMesh& m = ctx->meshes[*out];
	for (const auto& g : ps.groups) m.group_names.push_back(g.name);
	m.host_vertices.swap(v14);
	m.host_indices.swap(ps.idx);
	if (info) {
		info->n_verts = nv; info->n_faces = nf;
		info->n_groups = (uint32_t)ps.groups.size(); info->reserved = 0;
		info->first_drawn_face = ps.groups.empty() ? nf : ps.groups[0].first_face;
	}
	return AXR_OK;
}

int axr_load_obj_file(axr_ctx* ctx, const char* path, axr_mesh* out, axr_obj_info* info) {
	if (!ctx) return AXR_ERR_INVALID;
	if (!path || !out) return fail(ctx, AXR_ERR_INVALID, "axr_load_obj_file: null argument");
	FILE* fh = fopen(path, "rb");
	if (!fh) return fail(ctx, AXR_ERR_INVALID, "axr_load_obj_file: cannot open %s", path);
	std::string text;
	char buf[1 << 16];
	size_t n;
	while ((n = fread(buf, 1, sizeof buf, fh)) > 0) text.append(buf, n);
	fclose(fh);
	return axr_load_obj(ctx, text.data(), text.size(), out, info);
}

int axr_mesh_group_info(axr_ctx* ctx, axr_mesh mh, uint32_t group, char* name, size_t name_cap, uint64_t* first_face, uint64_t* face_count) {
	if (!ctx) return AXR_ERR_INVALID;
	if (!valid_mesh(ctx, mh)) return fail(ctx, AXR_ERR_INVALID, "axr_mesh_group_info: bad mesh handle %d", mh);
	const DeviceMesh& m = ctx->meshes[mh];
	if (group >= m.group_names.size()) return fail(ctx, AXR_ERR_INVALID, "axr_mesh_group_info: group %u of %zu (meshes loaded with axr_load_obj only)", group, m.group_names.size());
	if (name && name_cap) { strncpy(name, m.group_names[group].c_str(), name_cap - 1); name[name_cap - 1] = 0; }
	if (first_face) *first_face = m.group_first[group];
	if (face_count) *face_count = m.group_first[group + 1] - m.group_first[group];
	return AXR_OK;
}

int axr_mesh_read(axr_ctx* ctx, axr_mesh mh, float* vertices, uint32_t* indices) {
	if (!ctx) return AXR_ERR_INVALID;
	if (!valid_mesh(ctx, mh)) return fail(ctx, AXR_ERR_INVALID, "axr_mesh_read: bad mesh handle %d", mh);
	const DeviceMesh& m = ctx->meshes[mh];
	if (m.host_vertices.size() != m.n_verts * 14 || m.host_indices.size() != m.n_faces * 3)
		return fail(ctx, AXR_ERR_INVALID, "axr_mesh_read: the loader's arrays are kept for meshes loaded with axr_load_obj only");
	if (vertices && m.n_verts) memcpy(vertices, m.host_vertices.data(), m.host_vertices.size() * sizeof(float));
	if (indices && m.n_faces) memcpy(indices, m.host_indices.data(), m.host_indices.size() * sizeof(uint32_t));
	return AXR_OK;
}

int axr_parse_mtl(const char* text, size_t len, axr_mtl_entry* out, uint32_t cap, uint32_t* n_out) {
	if ((len && !text) || !n_out || (cap && !out)) return AXR_ERR_INVALID;
	std::vector<axr_obj::MtlEntry> v;
	axr_obj::parse_mtl(text, len, v);
	*n_out = (uint32_t)v.size();
	for (uint32_t i = 0; i < v.size() && i < cap; ++i) {
		memset(&out[i], 0, sizeof out[i]);
		strncpy(out[i].name, v[i].name.c_str(), sizeof out[i].name - 1);
		out[i].specular_exponent = v[i].specular_exponent;
		for (int k = 0; k < 5; ++k) {
			out[i].has_map[k] = v[i].has_map[k] ? 1 : 0;
			strncpy(out[i].map[k], v[i].map[k].c_str(), sizeof out[i].map[k] - 1);
		}
	}
	return AXR_OK;
}

int axr_generate_tangents(axr_ctx* ctx, const float* v8, uint64_t n_verts, const uint32_t* indices, uint64_t n_faces, float* out14) {
	if (!ctx) return AXR_ERR_INVALID;
	if ((n_verts && (!v8 || !out14)) || (n_faces && !indices)) return fail(ctx, AXR_ERR_INVALID, "axr_generate_tangents: null argument");
	if (n_verts >= (1ull << 31) - 1 || n_faces * 3 >= (1ull << 32)) return fail(ctx, AXR_ERR_CAPACITY, "axr_generate_tangents: mesh too large (vertex count for the 32-bit scan, corner ids)");
	for (uint64_t i = 0; i < n_faces * 3; ++i)
		if (indices[i] >= n_verts) return fail(ctx, AXR_ERR_INVALID, "axr_generate_tangents: index %u out of range", indices[i]);
	if (n_verts == 0) return AXR_OK;
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	cudaStream_t s = ctx->stream;
	float *d_v8 = nullptr, *d_out = nullptr;
	unsigned *d_idx = nullptr, *d_deg = nullptr, *d_start = nullptr, *d_cur = nullptr, *d_corners = nullptr;
	FaceTB* d_ftb = nullptr;
	void* d_tmp = nullptr;
	size_t tmp_bytes = 0;
	const size_t nf = n_faces ? n_faces : 1;
	auto cleanup = [&]() { cudaFree(d_v8); cudaFree(d_out); cudaFree(d_idx); cudaFree(d_deg); cudaFree(d_start); cudaFree(d_cur); cudaFree(d_corners); cudaFree(d_ftb); cudaFree(d_tmp); };
#define CUT(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(ctx, AXR_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } } while (0)
	CUT(cudaMalloc(&d_v8, n_verts * 8 * sizeof(float)));
	CUT(cudaMalloc(&d_out, n_verts * 14 * sizeof(float)));
	CUT(cudaMalloc(&d_idx, nf * 3 * sizeof(unsigned)));
	CUT(cudaMalloc(&d_deg, (n_verts + 1) * sizeof(unsigned)));
	CUT(cudaMalloc(&d_start, (n_verts + 1) * sizeof(unsigned)));
	CUT(cudaMalloc(&d_cur, n_verts * sizeof(unsigned)));
	CUT(cudaMalloc(&d_corners, nf * 3 * sizeof(unsigned)));
	CUT(cudaMalloc(&d_ftb, nf * sizeof(FaceTB)));
	CUT(cudaMemcpyAsync(d_v8, v8, n_verts * 8 * sizeof(float), cudaMemcpyHostToDevice, s));
	if (n_faces) CUT(cudaMemcpyAsync(d_idx, indices, n_faces * 3 * sizeof(unsigned), cudaMemcpyHostToDevice, s));
	CUT(cudaMemsetAsync(d_deg, 0, (n_verts + 1) * sizeof(unsigned), s));
	CUT(cudaMemsetAsync(d_cur, 0, n_verts * sizeof(unsigned), s));
	if (n_faces) k_tan_faces<<<(unsigned)((n_faces + 255) / 256), 256, 0, s>>>(d_v8, d_idx, n_faces, d_ftb, d_deg);
	CUT(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_deg, d_start, (int)(n_verts + 1), s));
	CUT(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1));
	CUT(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_deg, d_start, (int)(n_verts + 1), s));
	if (n_faces) k_tan_fill<<<(unsigned)((n_faces + 255) / 256), 256, 0, s>>>(d_idx, n_faces, d_start, d_cur, d_corners);
	k_tan_vertices<<<(unsigned)((n_verts + 255) / 256), 256, 0, s>>>(d_v8, n_verts, d_start, d_corners, d_ftb, d_out);
	CUT(cudaGetLastError());
	CUT(cudaMemcpyAsync(out14, d_out, n_verts * 14 * sizeof(float), cudaMemcpyDeviceToHost, s));
	CUT(cudaStreamSynchronize(s));
#undef CUT
	cleanup();
	return AXR_OK;
}

int axr_update_mesh_vertices(axr_ctx* ctx, axr_mesh mh, const float* vertices, uint64_t n_verts) {
	if (!ctx) return AXR_ERR_INVALID;
	if (!valid_mesh(ctx, mh) || !vertices) return fail(ctx, AXR_ERR_INVALID, "axr_update_mesh_vertices: bad mesh handle %d or null pointer", mh);
	DeviceMesh& m = ctx->meshes[mh];
	if (n_verts != m.n_verts) return fail(ctx, AXR_ERR_INVALID, "axr_update_mesh_vertices: %llu vertices, the mesh has %llu", (unsigned long long)n_verts, m.n_verts);
	if (n_verts == 0) return AXR_OK;
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	rc = sync_all(ctx);  // no draw in flight may still read the old vertices
	if (rc) return rc;
	float* raw = nullptr;
	CU(cudaMalloc(&raw, n_verts * 14 * sizeof(float)));
	cudaMemcpyAsync(raw, vertices, n_verts * 14 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
	k_split_vertices<<<(unsigned)((n_verts + 255) / 256), 256, 0, ctx->stream>>>(raw, n_verts, n_verts, m.pos, m.attr);
	const cudaError_t e = cudaStreamSynchronize(ctx->stream);
	cudaFree(raw);
	if (e != cudaSuccess) return fail(ctx, AXR_ERR_CUDA, "axr_update_mesh_vertices: %s", cudaGetErrorString(e));
	return AXR_OK;
}

int axr_free_mesh(axr_ctx* ctx, axr_mesh mh) {
	if (!ctx) return AXR_ERR_INVALID;
	if (!valid_mesh(ctx, mh)) return fail(ctx, AXR_ERR_INVALID, "axr_free_mesh: bad handle %d", mh);
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	rc = sync_all(ctx);
	if (rc) return rc;
	DeviceMesh& m = ctx->meshes[mh];
	cudaFree(m.pos); cudaFree(m.attr); cudaFree(m.idx); cudaFree(m.idx4); cudaFree(m.sv[0]); cudaFree(m.sv[1]); cudaFree(m.d_materials); cudaFree(m.d_group_first);
	m = DeviceMesh();
	return AXR_OK;
}

int axr_upload_texture(axr_ctx* ctx, const uint8_t* rgba, int w, int h, axr_tex* out) {
	if (!ctx) return AXR_ERR_INVALID;
	if (!rgba || !out || w <= 0 || h <= 0) return fail(ctx, AXR_ERR_INVALID, "axr_upload_texture: bad argument");
	CU(cudaSetDevice(ctx->device));
	if ((unsigned long long)w * (unsigned long long)h >= (1ull << 31)) return fail(ctx, AXR_ERR_CAPACITY, "axr_upload_texture: %dx%d exceeds 2^31 texels", w, h);
	DeviceTexture t;
	t.w = w; t.h = h;
	t.tiles_x = (w + 7) / 8;
	const size_t padded = (size_t)t.tiles_x * 8 * (size_t)((h + 3) / 4 * 4);
	uchar4* linear = nullptr;
	CU(cudaMalloc(&t.data, padded * 4));
	if (cudaMalloc(&linear, (size_t)w * h * 4) != cudaSuccess) { cudaFree(t.data); return fail(ctx, AXR_ERR_CUDA, "axr_upload_texture: out of device memory"); }
	cudaMemsetAsync(t.data, 0, padded * 4, ctx->stream);
	cudaMemcpyAsync(linear, rgba, (size_t)w * h * 4, cudaMemcpyHostToDevice, ctx->stream);
	k_tile_texture<<<dim3((w + 31) / 32, (h + 7) / 8), dim3(32, 8), 0, ctx->stream>>>(linear, w, h, t.tiles_x, t.data);
	const cudaError_t te = cudaStreamSynchronize(ctx->stream);
	cudaFree(linear);
	if (te != cudaSuccess) { cudaFree(t.data); return fail(ctx, AXR_ERR_CUDA, "axr_upload_texture: %s", cudaGetErrorString(te)); }
	t.live = true;
	size_t slot = ctx->textures.size();
	for (size_t i = 0; i < ctx->textures.size(); ++i) if (!ctx->textures[i].live) { slot = i; break; }
	if (slot == ctx->textures.size()) ctx->textures.push_back(t); else ctx->textures[slot] = t;
	*out = (axr_tex)slot;
	return AXR_OK;
}

int axr_free_texture(axr_ctx* ctx, axr_tex th) {
	if (!ctx) return AXR_ERR_INVALID;
	if (!valid_tex(ctx, th)) return fail(ctx, AXR_ERR_INVALID, "axr_free_texture: bad handle %d", th);
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	CU(cudaStreamSynchronize(ctx->stream));
	const uchar4* gone = ctx->textures[th].data;
	cudaFree(ctx->textures[th].data);
	ctx->textures[th] = DeviceTexture();
	for (auto& m : ctx->meshes)  // materials that referenced the texture lose it (a later draw then reports AXR_ERR_MATERIAL)
		if (m.live)
			for (auto& mat : m.materials)
				for (auto& tr : mat.tex)
					if (tr.data == gone) { tr = TexRef{nullptr, 0, 0, 0}; m.materials_dirty = true; }
	return AXR_OK;
}

int axr_set_material(axr_ctx* ctx, axr_mesh mh, uint32_t group, axr_tex diffuse, axr_tex bump, axr_tex metallic, axr_tex roughness,
                     axr_tex ao, float specular_exponent) {
	if (!ctx) return AXR_ERR_INVALID;
	if (!valid_mesh(ctx, mh)) return fail(ctx, AXR_ERR_INVALID, "axr_set_material: bad mesh handle %d", mh);
	DeviceMesh& m = ctx->meshes[mh];
	if (group >= m.materials.size()) return fail(ctx, AXR_ERR_INVALID, "axr_set_material: group %u of %zu", group, m.materials.size());
	const axr_tex th[5] = {diffuse, bump, metallic, roughness, ao};
	Material mat{};
	for (int i = 0; i < 5; ++i) {
		if (th[i] == AXR_NO_TEXTURE) { mat.tex[i] = TexRef{nullptr, 0, 0, 0}; continue; }
		if (!valid_tex(ctx, th[i])) return fail(ctx, AXR_ERR_INVALID, "axr_set_material: bad texture handle %d", th[i]);
		const DeviceTexture& t = ctx->textures[th[i]];
		mat.tex[i] = TexRef{t.data, t.w, t.h, t.tiles_x};
	}
	mat.specular_exponent = specular_exponent;
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	m.materials[group] = mat;
	m.materials_dirty = true;
	return AXR_OK;
}

int axr_set_uniforms(axr_ctx* ctx, const float view_proj[16], const float viewport[16], const float cam_pos[3]) {
	if (!ctx) return AXR_ERR_INVALID;
	if (!view_proj || !cam_pos) return fail(ctx, AXR_ERR_INVALID, "axr_set_uniforms: null argument");
	CU(cudaSetDevice(ctx->device));
	if (int rc = check_pending(ctx)) return rc;  // a draw that has to be redone is redone with the state it was issued with
	memcpy(ctx->view_proj, view_proj, sizeof ctx->view_proj);
	if (viewport) memcpy(ctx->viewport, viewport, sizeof ctx->viewport);
	memcpy(ctx->cam_pos, cam_pos, sizeof ctx->cam_pos);
	return AXR_OK;
}

int axr_set_shader(axr_ctx* ctx, int kind, const axr_shader_params* params, size_t params_size) {
	if (!ctx) return AXR_ERR_INVALID;
	if (kind != AXR_SHADER_FLAT && kind != AXR_SHADER_PHONG && kind != AXR_SHADER_PBR && kind != AXR_SHADER_CUTOUT && !shader_is_plugin(ctx, kind))
		return fail(ctx, AXR_ERR_UNSUPPORTED, "axr_set_shader: no device functor for shader kind %d (a further IShader is added with axr_load_shader_plugin)", kind);
	if (!params || params_size != sizeof(axr_shader_params)) return fail(ctx, AXR_ERR_INVALID, "axr_set_shader: bad params");
	CU(cudaSetDevice(ctx->device));
	if (int rc = check_pending(ctx)) return rc;
	ctx->shader_kind = kind;
	ctx->shader_params = *params;
	return AXR_OK;
}

int axr_set_shader_user(axr_ctx* ctx, const float* values, uint32_t n) {
	if (!ctx) return AXR_ERR_INVALID;
	if (n > 8 || (n && !values)) return fail(ctx, AXR_ERR_INVALID, "axr_set_shader_user: at most 8 floats");
	CU(cudaSetDevice(ctx->device));
	if (int rc = check_pending(ctx)) return rc;
	memset(ctx->shader_user, 0, sizeof ctx->shader_user);
	if (n) memcpy(ctx->shader_user, values, n * sizeof(float));
	return AXR_OK;
}

int axr_load_shader_plugin(axr_ctx* ctx, const char* path, int* kind_out) {
	if (!ctx) return AXR_ERR_INVALID;
	if (!path || !kind_out) return fail(ctx, AXR_ERR_INVALID, "axr_load_shader_plugin: null argument");
	for (size_t i = 0; i < ctx->plugins.size(); ++i)
		if (ctx->plugins[i].path == path) { *kind_out = AXR_SHADER_PLUGIN_BASE + (int)i; return AXR_OK; }
	void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
	if (!h) return fail(ctx, AXR_ERR_INVALID, "axr_load_shader_plugin: %s", dlerror());
	auto layout = (unsigned long long (*)(void))dlsym(h, "axr_shader_plugin_layout");
	auto discards = (int (*)(void))dlsym(h, "axr_shader_plugin_discards");
	auto textures = (unsigned (*)(void))dlsym(h, "axr_shader_plugin_textures");
	ShaderPlugin p;
	p.launch = (decltype(p.launch))dlsym(h, "axr_shader_plugin_launch");
	if (!layout || !discards || !textures || !p.launch) {
		dlclose(h);
		return fail(ctx, AXR_ERR_INVALID, "axr_load_shader_plugin: %s does not export the AXR_SHADER_PLUGIN entry points", path);
	}
	if (layout() != plugin_layout_hash()) {
		dlclose(h);
		return fail(ctx, AXR_ERR_UNSUPPORTED, "axr_load_shader_plugin: %s was built against other kernel headers than this library (rebuild it)", path);
	}
	p.handle = h; p.discards = discards() != 0; p.textures = textures(); p.path = path;
	ctx->plugins.push_back(p);
	*kind_out = AXR_SHADER_PLUGIN_BASE + (int)ctx->plugins.size() - 1;
	return AXR_OK;
}

int axr_set_sampler(axr_ctx* ctx, int sampler) {
	if (!ctx) return AXR_ERR_INVALID;
	if (sampler != AXR_SAMPLER_NEAREST && sampler != AXR_SAMPLER_BILINEAR) return fail(ctx, AXR_ERR_INVALID, "axr_set_sampler: %d", sampler);
	CU(cudaSetDevice(ctx->device));
	if (int rc = check_pending(ctx)) return rc;
	ctx->sampler = sampler;
	return AXR_OK;
}

int axr_clear(axr_ctx* ctx, uint32_t packed_argb, float depth) {
	if (!ctx) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	const size_t first = (size_t)ctx->fp.y_lo * ctx->fp.W, n = (size_t)(ctx->fp.y_hi - ctx->fp.y_lo) * ctx->fp.W;
	k_clear<<<grid_for(n, 256), 256, 0, ctx->stream>>>(ctx->out_color, ctx->out_depth, packed_argb, depth, first, n);
	CU(cudaGetLastError());
	ctx->depth_fresh = depth == INFINITY;
	ctx->fresh_target = ctx->out_depth;
	return AXR_OK;
}

static int upload_framebuffer(axr_ctx* ctx, const uint8_t* bgra, const float* depth, bool wait) {
	if (!ctx) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	const size_t first = (size_t)ctx->fp.y_lo * ctx->fp.W, n = (size_t)(ctx->fp.y_hi - ctx->fp.y_lo) * ctx->fp.W;
	if (bgra) CU(cudaMemcpyAsync(ctx->out_color + first, bgra + first * 4, n * 4, cudaMemcpyHostToDevice, ctx->stream));
	if (depth) {
		CU(cudaMemcpyAsync(ctx->out_depth + first, depth + first, n * 4, cudaMemcpyHostToDevice, ctx->stream));
		ctx->depth_fresh = false;
	}
	if (wait) CU(cudaStreamSynchronize(ctx->stream));
	return AXR_OK;
}

int axr_upload_framebuffer(axr_ctx* ctx, const uint8_t* bgra, const float* depth) { return upload_framebuffer(ctx, bgra, depth, true); }
int axr_upload_framebuffer_async(axr_ctx* ctx, const uint8_t* bgra, const float* depth) { return upload_framebuffer(ctx, bgra, depth, false); }

int axr_resolve(axr_ctx* ctx, uint8_t* bgra_out, float* depth_out) {
	if (!ctx) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	const size_t first = (size_t)ctx->fp.y_lo * ctx->fp.W, n = (size_t)(ctx->fp.y_hi - ctx->fp.y_lo) * ctx->fp.W;
	if (bgra_out) CU(cudaMemcpyAsync(bgra_out + first * 4, ctx->out_color + first, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (depth_out) CU(cudaMemcpyAsync(depth_out + first, ctx->out_depth + first, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return AXR_OK;
}

int axr_draw_mesh(axr_ctx* ctx, axr_mesh mh, const float model[16]) {
	if (!ctx) return AXR_ERR_INVALID;
	if (!valid_mesh(ctx, mh) || !model) return fail(ctx, AXR_ERR_INVALID, "axr_draw_mesh: bad mesh handle %d or null model", mh);
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	return draw(ctx, mh, model);
}

// Device-side alias of a host range, page-locking it on first use. Returns nullptr when the range cannot be mapped.
// Ranges this context registered are reused only when they contain [p, p + bytes); a registration that starts inside the range but
// is too small (a freed buffer whose address was reused by a larger framebuffer) is dropped and taken again. Memory the caller
// pinned itself (axr_host_alloc, cudaHostAlloc) is taken at its word: the allocation has to cover the framebuffer.
static void* map_host_range(axr_ctx* ctx, void* p, size_t bytes) {
	void* d = nullptr;
	char* const lo = (char*)p;
	char* const hi = lo + bytes;
	for (size_t i = 0; i < ctx->registered.size();) {
		char* const rlo = (char*)ctx->registered[i].first;
		char* const rhi = rlo + ctx->registered[i].second;
		if (lo >= rlo && hi <= rhi) return cudaHostGetDevicePointer(&d, p, 0) == cudaSuccess ? d : nullptr;
		if (lo < rhi && hi > rlo) {  // overlaps without containing: stale
			cudaHostUnregister(ctx->registered[i].first);
			cudaGetLastError();
			ctx->registered.erase(ctx->registered.begin() + (long)i);
			continue;
		}
		++i;
	}
	cudaPointerAttributes attr;  // (asking for the device pointer of pageable memory is an error; asking what the memory is, is not)
	if (cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost)
		return cudaHostGetDevicePointer(&d, p, 0) == cudaSuccess ? d : nullptr;  // pinned by the caller
	cudaGetLastError();
	if (cudaHostRegister(p, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	ctx->registered.emplace_back(p, bytes);
	if (cudaHostGetDevicePointer(&d, p, 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	return d;
}

int axr_host_release(axr_ctx* ctx, void* host_ptr) {
	if (!ctx) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	if (int rc = check_pending(ctx)) return rc;
	if (int rc = sync_all(ctx)) return rc;
	for (size_t i = 0; i < ctx->registered.size(); ++i) {
		char* const rlo = (char*)ctx->registered[i].first;
		if ((char*)host_ptr >= rlo && (char*)host_ptr < rlo + ctx->registered[i].second) {
			cudaHostUnregister(ctx->registered[i].first);
			cudaGetLastError();
			ctx->registered.erase(ctx->registered.begin() + (long)i);
			return AXR_OK;
		}
	}
	return AXR_OK;  // never registered (pinned by the caller, or not mappable): nothing to release
}

int axr_draw_mesh_host(axr_ctx* ctx, axr_mesh mh, const float model[16], uint8_t* bgra, float* depth) {
	if (!ctx) return AXR_ERR_INVALID;
	if (!valid_mesh(ctx, mh) || !model || !bgra || !depth) return fail(ctx, AXR_ERR_INVALID, "axr_draw_mesh_host: bad mesh handle %d or null pointer", mh);
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	const size_t npx = (size_t)ctx->fp.W * ctx->fp.H;
	const size_t first = (size_t)ctx->fp.y_lo * ctx->fp.W, n = (size_t)(ctx->fp.y_hi - ctx->fp.y_lo) * ctx->fp.W;
	void* dc = map_host_range(ctx, bgra, npx * 4);
	void* dd = dc ? map_host_range(ctx, depth, npx * 4) : nullptr;
	if (!dc || !dd) {  // not mappable: the plain round trip
		rc = upload_framebuffer(ctx, bgra, depth, false);
		if (!rc) rc = draw(ctx, mh, model);
		if (!rc) rc = axr_resolve(ctx, bgra, depth);
		return rc;
	}
	// host depth -> device copy for the merge test, in chunks of whole GPU tile rows on the upload stream; the geometry stages do
	// not wait for it, and the tile kernel of chunk b only waits for chunk b (launch_tile)
	(void)first; (void)n;
	const int ty_lo = ctx->fp.ty_lo, ty_hi = ctx->fp.ty_hi;
	int chunks = ctx->shader_kind == AXR_SHADER_CUTOUT ? 1 : (ty_hi - ty_lo) / 16;  // >= 16 tile rows (512 px) per chunk (C3 e2e: 1.22 / 1.04 / 1.06 / 1.09 ms with 1 / 2 / 4 / 8 chunks); peeled draws re-read the copy
	if (chunks > axr_ctx::MAX_HOST_CHUNKS) chunks = axr_ctx::MAX_HOST_CHUNKS;
	if (chunks < 1 || ctx->host_depth_zero_copy) chunks = 1;
	CU(cudaEventRecord(ctx->depth_free, ctx->stream));
	CU(cudaStreamWaitEvent(ctx->up_stream, ctx->depth_free, 0));
	for (int b = 0; b <= chunks; ++b) ctx->chunk_ty[b] = ty_lo + (int)((long long)(ty_hi - ty_lo) * b / chunks);
	for (int b = 0; b < chunks; ++b) {
		int r0 = ctx->chunk_ty[b] * GT, r1 = ctx->chunk_ty[b + 1] * GT;
		if (r0 < ctx->fp.y_lo) r0 = ctx->fp.y_lo;
		if (r1 > ctx->fp.y_hi) r1 = ctx->fp.y_hi;
		if (r1 > r0) {
			const size_t off = (size_t)r0 * ctx->fp.W, cnt = (size_t)(r1 - r0) * ctx->fp.W;
			if (!ctx->host_depth_zero_copy) CU(cudaMemcpyAsync(ctx->depth + off, depth + off, cnt * 4, cudaMemcpyHostToDevice, ctx->up_stream));
		}
		CU(cudaEventRecord(ctx->up_done[b], ctx->up_stream));
	}
	unsigned* save_c = ctx->out_color; float* save_d = ctx->out_depth;
	ctx->out_color = (unsigned*)dc; ctx->out_depth = (float*)dd;
	ctx->depth_read_override = ctx->host_depth_zero_copy ? (const float*)dd : ctx->depth;
	ctx->host_chunks = chunks;
	rc = draw(ctx, mh, model);
	if (!rc) rc = check_pending(ctx);  // a bin overflow re-issues the draw while the host pointers are still installed
	ctx->host_chunks = 0;
	ctx->out_color = save_c; ctx->out_depth = save_d; ctx->depth_read_override = nullptr;
	if (rc) return rc;
	return sync_all(ctx);  // complete on return: the kernel's stores have landed in the host arrays
}

int axr_sync(axr_ctx* ctx) {
	if (!ctx) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	CU(cudaStreamSynchronize(ctx->stream));
	return AXR_OK;
}

int axr_get_stats(axr_ctx* ctx, axr_stats* out) {
	if (!ctx || !out) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	*out = ctx->stats;
	return AXR_OK;
}

int axr_set_profiling(axr_ctx* ctx, int enabled) {
	if (!ctx) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	CU(cudaStreamSynchronize(ctx->geom_stream));
	CU(cudaStreamSynchronize(ctx->stream));
	for (cudaEvent_t e : ctx->prof_events) ctx->prof_pool.push_back(e);
	ctx->prof_events.clear();
	ctx->profiling = enabled != 0;
	return AXR_OK;
}

int axr_get_kernel_times(axr_ctx* ctx, float ms_out[AXR_NUM_STAGES], uint64_t* draws_out) {
	if (!ctx || !ms_out) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	CU(cudaStreamSynchronize(ctx->stream));
	const size_t per = 2 * AXR_NUM_STAGES;  // (begin, end) per stage, each pair recorded on the stream its kernel runs on
	const size_t draws = ctx->prof_events.size() / per;
	for (int k = 0; k < AXR_NUM_STAGES; ++k) ms_out[k] = 0.f;
	for (size_t d = 0; d < draws; ++d)
		for (int k = 0; k < AXR_NUM_STAGES; ++k) {
			float ms = 0.f;
			CU(cudaEventElapsedTime(&ms, ctx->prof_events[d * per + 2 * k], ctx->prof_events[d * per + 2 * k + 1]));
			ms_out[k] += ms;
		}
	for (cudaEvent_t e : ctx->prof_events) ctx->prof_pool.push_back(e);
	ctx->prof_events.clear();
	if (draws_out) *draws_out = draws;
	return AXR_OK;
}

int axr_measure_fp32_issue(axr_ctx* ctx, double* fmul_fadd_winst_per_s, double* ffma_winst_per_s) {
	if (!ctx) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	rc = sync_all(ctx);
	if (rc) return rc;
	const int ctas = 148 * 8, threads = 256, iters = 2048;
	float* buf = nullptr;
	CU(cudaMalloc(&buf, (size_t)ctas * threads * sizeof(float)));
	cudaEvent_t e0, e1;
	CU(cudaEventCreate(&e0));
	CU(cudaEventCreate(&e1));
	double res[2] = {0, 0};
	for (int mode = 0; mode < 2; ++mode) {
		float best = 1e30f;
		for (int rep = 0; rep < 4; ++rep) {  // first repetition warms up
			cudaEventRecord(e0, ctx->stream);
			if (mode == 0) k_fp32_peak<0><<<ctas, threads, 0, ctx->stream>>>(buf, 1.0000001f, 1e-7f, iters);
			else k_fp32_peak<1><<<ctas, threads, 0, ctx->stream>>>(buf, 1.0000001f, 1e-7f, iters);
			cudaEventRecord(e1, ctx->stream);
			cudaEventSynchronize(e1);
			float ms = 0.f;
			cudaEventElapsedTime(&ms, e0, e1);
			if (rep && ms < best) best = ms;
		}
		const double warp_inst = (double)ctas * (threads / 32) * (double)iters * 16.0 * 8.0 * (mode == 0 ? 2.0 : 1.0);
		res[mode] = warp_inst / (best * 1e-3);
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(buf);
	CU(cudaGetLastError());
	if (fmul_fadd_winst_per_s) *fmul_fadd_winst_per_s = res[0];
	if (ffma_winst_per_s) *ffma_winst_per_s = res[1];
	return AXR_OK;
}

int axr_measure_gather(axr_ctx* ctx, double* sectors_per_s) {
	if (!ctx || !sectors_per_s) return AXR_ERR_INVALID;
	const int span = getenv("AXR_GATHER_SPAN") ? atoi(getenv("AXR_GATHER_SPAN")) : 1;  // measurement knob: 2 / 4 consecutive sectors per position
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	rc = sync_all(ctx);
	if (rc) return rc;
	const unsigned long long n_sectors = 1ull << 25;  // 1 GiB, 8x the L2
	const int ctas = 148 * 8, threads = 256, steps = 64;
	constexpr int ILP = 8;
	uint4* buf = nullptr;
	unsigned* out = nullptr;
	CU(cudaMalloc(&buf, n_sectors * 32));
	if (cudaMalloc(&out, (size_t)ctas * threads * 4) != cudaSuccess) { cudaFree(buf); return fail(ctx, AXR_ERR_CUDA, "axr_measure_gather: out of device memory"); }
	cudaMemsetAsync(buf, 0x5a, n_sectors * 32, ctx->stream);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	float best = 1e30f;
	for (int rep = 0; rep < 4; ++rep) {  // first repetition warms up
		cudaEventRecord(e0, ctx->stream);
		if (span == 4) k_gather_peak<ILP / 2, 4><<<ctas, threads, 0, ctx->stream>>>(buf, n_sectors, steps, out);
		else if (span == 2) k_gather_peak<ILP / 2, 2><<<ctas, threads, 0, ctx->stream>>>(buf, n_sectors, steps, out);
		else k_gather_peak<ILP, 1><<<ctas, threads, 0, ctx->stream>>>(buf, n_sectors, steps, out);
		cudaEventRecord(e1, ctx->stream);
		cudaEventSynchronize(e1);
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		if (rep && ms < best) best = ms;
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(buf);
	cudaFree(out);
	CU(cudaGetLastError());
	*sectors_per_s = (double)ctas * threads * steps * (span == 1 ? ILP : (ILP / 2) * span) / (best * 1e-3);
	return AXR_OK;
}

void* axr_host_alloc(size_t bytes) {
	void* p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) return nullptr;
	return p;
}
void axr_host_free(void* p) {
	if (p) cudaFreeHost(p);
}

void* axr_stream(axr_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int axr_framebuffer_device(axr_ctx* ctx, void** bgra_dev, void** depth_dev) {
	if (!ctx) return AXR_ERR_INVALID;
	if (bgra_dev) *bgra_dev = ctx->color;
	if (depth_dev) *depth_dev = ctx->depth;
	ctx->depth_external = true;  // whoever holds the pointers may write the planes: no assumptions about their contents any more
	return AXR_OK;
}

int axr_set_output(axr_ctx* ctx, void* bgra_dev, void* depth_dev) {
	if (!ctx) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	if ((bgra_dev == nullptr) != (depth_dev == nullptr)) return fail(ctx, AXR_ERR_INVALID, "axr_set_output: pass both pointers or neither");
	ctx->out_color = bgra_dev ? (unsigned*)bgra_dev : ctx->color;
	ctx->out_depth = depth_dev ? (float*)depth_dev : ctx->depth;
	return AXR_OK;
}

int axr_dirty_map_entries(const axr_ctx* ctx) { return ctx ? ctx->fp.ntx * ctx->fp.nty : 0; }

int axr_set_dirty_map(axr_ctx* ctx, void* dirty_dev) {
	if (!ctx) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	if (int rc = check_pending(ctx)) return rc;
	ctx->dirty_map = (unsigned*)dirty_dev;
	return AXR_OK;
}

int axr_clear_dirty_tiles(axr_ctx* ctx, void* bgra_dev, void* depth_dev, void* dirty_dev, int count, uint32_t packed_argb, float depth, void* stream) {
	if (!ctx) return AXR_ERR_INVALID;
	if (!bgra_dev || !depth_dev || !dirty_dev || count <= 0 || count > 65535) return fail(ctx, AXR_ERR_INVALID, "axr_clear_dirty_tiles: bad argument");
	CU(cudaSetDevice(ctx->device));
	k_clear_dirty_tiles<<<dim3(ctx->fp.ntx, ctx->fp.nty, count), 256, 0, stream ? (cudaStream_t)stream : ctx->stream>>>(
		(unsigned*)bgra_dev, (float*)depth_dev, (unsigned*)dirty_dev, ctx->fp.W, ctx->fp.H, ctx->fp.ntx, GT, packed_argb, depth);
	CU(cudaGetLastError());
	return AXR_OK;
}

int axr_set_output_fill(axr_ctx* ctx, int enabled, uint32_t packed_argb, float depth) {
	if (!ctx) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	if (int rc = check_pending(ctx)) return rc;
	ctx->fill = enabled != 0; ctx->fill_color = packed_argb; ctx->fill_depth = depth;
	return AXR_OK;
}

int axr_set_output_rows(axr_ctx* ctx, int enabled) {
	if (!ctx) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	if (int rc = check_pending(ctx)) return rc;
	ctx->out_rows = enabled != 0;
	return AXR_OK;
}

int axr_clear_stale_tiles(axr_ctx* ctx, void* bgra_dev, void* depth_dev, void* dirty_prev_dev, void* dirty_now_dev, int count, uint32_t packed_argb, float depth,
                          void* stream) {
	if (!ctx) return AXR_ERR_INVALID;
	if (!bgra_dev || !depth_dev || !dirty_prev_dev || !dirty_now_dev || count <= 0 || count > 65535) return fail(ctx, AXR_ERR_INVALID, "axr_clear_stale_tiles: bad argument");
	CU(cudaSetDevice(ctx->device));
	cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
	const unsigned n_entries = (unsigned)n_tiles(ctx) * (unsigned)count;
	if (ctx->stale_cap < n_entries) {  // list of stale (target, tile) entries + its two counters; grown on first use
		CU(cudaDeviceSynchronize());
		cudaFree(ctx->stale_list);
		ctx->stale_list = nullptr; ctx->stale_cap = 0;
		CU(cudaMalloc(&ctx->stale_list, ((size_t)n_entries + 2) * sizeof(unsigned)));
		CU(cudaMemsetAsync(ctx->stale_list, 0, 2 * sizeof(unsigned), s));
		ctx->stale_cap = n_entries;
	}
	unsigned* counters = ctx->stale_list;  // [0] entries listed, [1] exit ticket
	k_find_stale_tiles<<<(n_entries + 255) / 256, 256, 0, s>>>((unsigned*)dirty_prev_dev, (const unsigned*)dirty_now_dev, n_entries, ctx->stale_list + 2, counters);
	k_clear_listed_tiles<<<148, 256, 0, s>>>((unsigned*)bgra_dev, (float*)depth_dev, ctx->stale_list + 2, counters, ctx->fp.W, ctx->fp.H, ctx->fp.ntx, ctx->fp.nty, GT,
	                                         packed_argb, depth);
	CU(cudaGetLastError());
	return AXR_OK;
}

int axr_set_overlap(axr_ctx* ctx, int enabled) {
	if (!ctx) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	rc = sync_all(ctx);
	if (rc) return rc;
	ctx->overlap = enabled != 0;
	return AXR_OK;
}

int axr_set_color_math(axr_ctx* ctx, int mode) {
	if (!ctx) return AXR_ERR_INVALID;
	if (mode != AXR_COLOR_EXACT && mode != AXR_COLOR_FAST) return fail(ctx, AXR_ERR_INVALID, "axr_set_color_math: %d", mode);
	CU(cudaSetDevice(ctx->device));
	if (int rc = check_pending(ctx)) return rc;
	ctx->color_fast = mode == AXR_COLOR_FAST;
	return AXR_OK;
}

int axr_set_depth_read(axr_ctx* ctx, int enabled) {
	if (!ctx) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	int rc = check_pending(ctx);
	if (rc) return rc;
	ctx->read_depth = enabled ? 1 : 0;
	return AXR_OK;
}

int axr_alloc_shared(axr_ctx* ctx, size_t bytes, void** dev_ptr_out, void* handle64_out) {
	if (!ctx || !dev_ptr_out || !handle64_out || bytes == 0) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	void* p = nullptr;
	CU(cudaMalloc(&p, bytes));
	CU(cudaMemsetAsync(p, 0, bytes, ctx->stream));  // dirty maps start out clean
	CU(cudaStreamSynchronize(ctx->stream));
	cudaError_t e = cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64_out, p);
	if (e != cudaSuccess) { cudaFree(p); return fail(ctx, AXR_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
	ctx->shared_allocs.push_back(p);
	*dev_ptr_out = p;
	return AXR_OK;
}

int axr_free_shared(axr_ctx* ctx, void* dev_ptr) {
	if (!ctx || !dev_ptr) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	for (size_t i = 0; i < ctx->shared_allocs.size(); ++i)
		if (ctx->shared_allocs[i] == dev_ptr) {
			CU(cudaDeviceSynchronize());
			CU(cudaFree(dev_ptr));
			ctx->shared_allocs.erase(ctx->shared_allocs.begin() + i);
			return AXR_OK;
		}
	return fail(ctx, AXR_ERR_INVALID, "axr_free_shared: pointer was not allocated by this context");
}

int axr_framebuffer_ipc(axr_ctx* ctx, void* color_handle64, void* depth_handle64) {
	if (!ctx || !color_handle64 || !depth_handle64) return AXR_ERR_INVALID;
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	CU(cudaSetDevice(ctx->device));
	CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)color_handle64, ctx->color));
	CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)depth_handle64, ctx->depth));
	ctx->depth_external = true;
	return AXR_OK;
}

int axr_open_ipc(axr_ctx* ctx, const void* handle64, void** dev_ptr_out) {
	if (!ctx || !handle64 || !dev_ptr_out) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, sizeof h);
	CU(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
	ctx->ipc_opened.push_back(*dev_ptr_out);
	return AXR_OK;
}

int axr_close_ipc(axr_ctx* ctx, void* dev_ptr) {
	if (!ctx || !dev_ptr) return AXR_ERR_INVALID;
	CU(cudaSetDevice(ctx->device));
	for (size_t i = 0; i < ctx->ipc_opened.size(); ++i)
		if (ctx->ipc_opened[i] == dev_ptr) {
			CU(cudaStreamSynchronize(ctx->stream));
			CU(cudaIpcCloseMemHandle(dev_ptr));
			ctx->ipc_opened.erase(ctx->ipc_opened.begin() + i);
			return AXR_OK;
		}
	return fail(ctx, AXR_ERR_INVALID, "axr_close_ipc: pointer was not opened by this context");
}

}  // extern "C"
