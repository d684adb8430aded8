// CUDA kernels of the B200 tiled-raster path (sm_100a). One draw =
//   k_vertex_xform   per unique vertex : mvp transform, exact + "safely outside" clip codes, perspective divide -> 16 B screen
//                                        record (also zeroes the draw's counters)
//   k_setup_raster   per input face    : 3 index loads + 3 record gathers; all-inside -> back-face test; provably outside ->
//                                        nothing; else McGuire clip (slow path). Triangles whose pixel box is small (<= 32 or
//                                        64 px, per draw) are rasterised right here by their own thread: exact coverage per
//                                        pixel, 64-bit RED.MIN of (orderable(z) << 32 | face*8+sub) into the visibility buffer,
//                                        warp-aggregated tile flags. Larger ones: warp-aggregated append of a 40 B record +
//                                        per-tile count
//   k_scan_tiles     one CTA           : block-wide exclusive prefix sum of the per-tile counts, folds the striped counters,
//                                        publishes the draw status to mapped host memory
//   k_bin_scatter    per record        : scatter record ids into the per-tile lists (order-free: the key carries the ordinal)
//   k_tile_shade<S,M> one CTA per 32x32 tile : stage the tile's keys in shared memory (and hand the global ones back empty), raster
//                                        the binned records (lane-per-record setup, shared atomicMin), then per visible pixel:
//                                        gather indices -> screen records / positions / attributes / framebuffer depth in one
//                                        batch -> setup -> barycentrics -> depth test -> IShader::vertex x3 accumulated into
//                                        varyings -> IShader::fragment (textures from HBM) -> BGRA8 + f32 store
// A shader whose fragment() may discard (Shader::DISCARDS) makes visibility depend on shading; such a draw is depth-peeled:
// the same kernels run in passes, every pass only accepts keys above the pixel's `floor` (the key of the fragment discarded
// there in the previous pass), and the host repeats until no winner was discarded (axr_api.cu: draw_peeled).
// Nothing here is a dense contraction, so there is no tensor-core work; the path is gather-, latency- and FP32-issue bound.
#pragma once
#include <cooperative_groups.h>

#include "axr_shaders.cuh"

namespace cg = cooperative_groups;

namespace axr {

#ifndef AXR_VERTEX_PER_THREAD
#define AXR_VERTEX_PER_THREAD 2  // vertices per thread of k_vertex_xform (two loads in flight per thread: C3 34.9 -> 32.0 us; four: the same)
#endif
constexpr int GT = 32;            // GPU tile edge in pixels (a multiple of REF_TILE; keeps rows 128 B wide for the resolve)
constexpr int GT_PIX = GT * GT;
// Triangles whose pixel box is at most small_dim on a side and small_area pixels are rasterised by their own thread in
// k_setup_raster; the rest go through the tile bins. The limits are per draw (SetupOut): a throughput-bound draw of millions of
// faces prefers the larger pair (the binned path costs more per triangle up to ~64 px: C3 at 8K 1.31 -> 1.07 ms), a draw of a
// few thousand faces is latency-bound and prefers the smaller one (a lone thread walking 64 pixels is the critical path:
// C1 0.111 -> 0.122 ms). Results are identical either way.
constexpr int SMALL_DIM_LAT = 8, SMALL_AREA_LAT = 32, SMALL_DIM_TPUT = 12, SMALL_AREA_TPUT = 64;
constexpr unsigned long long SMALL_TPUT_MIN_FACES = 500000ull;
// 8 warps per CTA, 4 CTAs per SM (64 registers). 4 warps x 8 CTAs is the same occupancy in finer scheduling units and was measured:
// C3 201 -> 198 us, but C1 58 -> 65 us and C2 20 -> 25 us (half as many warps walk a tile's bins / pixels); 4 x 9 or 4 x 10 spill
// (C3: 219 / 231 us).
#ifndef AXR_TILE_THREADS
#define AXR_TILE_THREADS 256
#endif
constexpr int TILE_THREADS = AXR_TILE_THREADS;
static_assert(TILE_THREADS % 32 == 0 && GT_PIX % TILE_THREADS == 0, "whole warps, whole pixel batches");

// 40-byte setup record of a triangle that goes through the tile bins
struct __align__(8) TriRecord {
	float x0, y0, x1, y1, x2, y2, z0, z1, z2;
	unsigned ordinal;
};

// Device-side draw status / counters (copied to pinned host memory after the scan)
struct DrawStatus {
	unsigned long long clipped_faces, triangles, small_triangles, binned_triangles;  // binned_triangles = records wanted
	unsigned long long bin_refs;                                                      // refs wanted
	unsigned overflow;                                                                // OVF_* bits
	unsigned pad;
	// per-warp partial counters are spread over STAT_STRIPES slots (no single hot address); the last CTA of k_setup_clipped folds them
	unsigned stripes[512][4];  // [stripe][STRIPE_*]: triangles and small triangles adjacent (one 64-bit reduction adds both)
};
constexpr int STAT_STRIPES = 512;
enum { STRIPE_TRIS = 0, STRIPE_SMALL = 1, STRIPE_BINNED = 2, STRIPE_SPARE = 3 };

struct FrameParams {
	int W, H;
	int y_lo, y_hi;      // band rows [y_lo, y_hi)
	int ntx, nty;        // GPU tiles over the full frame
	int ty_lo, ty_hi;    // tile rows touched by the band
};

struct MeshView {
	const float4* pos;        // x,y,z,1
	const float4* attr;       // AXR_ATTR_PLANES: three planes of n_plane float4 (uv + n.xy | n.z + t | b + pad); else 48 B VAttr records
	unsigned long long n_plane;
	const uint4* idx4;        // AXR_IDX_PAD: (i0, i1, i2, -) per face for the shading stage's single 16 B index gather
	const unsigned* idx;      // 3 per face
	unsigned long long n_verts, n_faces;
	const Material* materials;
	Material material0;                     // copy of materials[0]: single-group meshes read it from the kernel parameters (constant bank)
	const unsigned long long* group_first;  // n_groups + 1 entries (ascending), only read when n_groups > 1
	int n_groups;
	unsigned first_face;                    // faces before it belong to no material group and are not drawn
};

// ------------------------------------------------------------------------------------------------ utility kernels
__global__ void k_fill_u64(unsigned long long* p, unsigned long long v, size_t n) {
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	size_t stride = (size_t)gridDim.x * blockDim.x;
	for (; i < n; i += stride) p[i] = v;
}
__global__ void k_fill_u32(unsigned* p, unsigned v, size_t n) {
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	size_t stride = (size_t)gridDim.x * blockDim.x;
	for (; i < n; i += stride) p[i] = v;
}
// Framebuffer::clearColor / clearDepth (reference src/framebuffer.cpp:26-42) over rows [y_lo, y_hi)
__global__ void k_clear(unsigned* color, float* depth, unsigned packed, float z, size_t first, size_t n) {
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	size_t stride = (size_t)gridDim.x * blockDim.x;
	for (; i < n; i += stride) { color[first + i] = packed; depth[first + i] = z; }
}
// Layout knobs of the shading stage's gathers (A/B: tools/build_variants.py). A warp-wide 16 B gather is served a quarter-warp at a
// time, one L1 wavefront per distinct 128 B line; neighbouring pixels reference neighbouring vertices, so float4 PLANES (8 vertices
// per line) need about half the wavefronts of 48 B records (2.7 vertices per line).
#ifndef AXR_ATTR_PLANES
#define AXR_ATTR_PLANES 1
#endif
#ifndef AXR_IDX_PAD
#define AXR_IDX_PAD 0
#endif
// Framebuffer clear restricted to the tiles a draw stored into (dirty map written by k_tile_shade): `count` targets of W*H pixels
// laid out back to back (colour planes, depth planes, dirty maps each contiguous); the flags are reset. One CTA per tile and target.
__global__ void __launch_bounds__(256) k_clear_dirty_tiles(unsigned* color, float* depth, unsigned* dirty, int W, int H, int ntx, int tile_px,
                                                           unsigned packed, float z) {
	const size_t npx = (size_t)W * H;
	const int tile = blockIdx.y * ntx + blockIdx.x;
	unsigned* flag = dirty + (size_t)blockIdx.z * ((size_t)gridDim.x * gridDim.y) + tile;
	if (*flag == 0u) return;
	__syncthreads();
	if (threadIdx.x == 0) *flag = 0u;
	const int x0 = blockIdx.x * tile_px, y0 = blockIdx.y * tile_px;
	for (int p = threadIdx.x; p < tile_px * tile_px; p += blockDim.x) {
		const int px = x0 + p % tile_px, py = y0 + p / tile_px;
		if (px < W && py < H) {
			const size_t gi = (size_t)blockIdx.z * npx + (size_t)py * W + px;
			color[gi] = packed; depth[gi] = z;
		}
	}
}
// Companion of axr_set_output_fill: a tile that was stored into the previous time the target was used (`prev`) and not this time (`now`)
// still holds the old frame and has to be cleared; `prev` is handed back all zero (it is the next frame's `now`). Two small kernels:
// one THREAD per (target, tile) looks at the two flags and lists the stale tiles (usually none: a CTA per tile just to look would put
// tens of thousands of CTAs on GPU 0 every frame, beside the view it is rendering), then a fixed grid clears the listed ones.
__global__ void __launch_bounds__(256) k_find_stale_tiles(unsigned* prev, const unsigned* now, unsigned n_entries, unsigned* list, unsigned* n_list) {
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i == 0) n_list[1] = 0u;  // (the clear kernel's exit ticket, see below)
	if (i >= n_entries) return;
	const unsigned was = prev[i];
	if (was == 0u) return;
	prev[i] = 0u;
	if (now[i] == 0u) list[atomicAdd(n_list, 1u)] = i;
}
__global__ void __launch_bounds__(256) k_clear_listed_tiles(unsigned* color, float* depth, const unsigned* list, unsigned* n_list, int W, int H, int ntx, int nty,
                                                            int tile_px, unsigned packed, float z) {
	const size_t npx = (size_t)W * H;
	const unsigned n = *n_list;
	for (unsigned e = blockIdx.x; e < n; e += gridDim.x) {
		const unsigned i = list[e], target = i / (unsigned)(ntx * nty), tile = i % (unsigned)(ntx * nty);
		const int x0 = (int)(tile % (unsigned)ntx) * tile_px, y0 = (int)(tile / (unsigned)ntx) * tile_px;
		for (int p = threadIdx.x; p < tile_px * tile_px; p += blockDim.x) {
			const int px = x0 + p % tile_px, py = y0 + p / tile_px;
			if (px < W && py < H) {
				const size_t gi = (size_t)target * npx + (size_t)py * W + px;
				color[gi] = packed; depth[gi] = z;
			}
		}
	}
	// the last CTA to leave resets the list for the next call
	__syncthreads();
	if (threadIdx.x == 0 && atomicAdd(n_list + 1, 1u) == gridDim.x - 1) n_list[0] = 0u;
}
// AoS AR::Vertex (56 B) -> position float4 + attributes (done once at mesh upload)
__global__ void k_split_vertices(const float* __restrict__ raw, unsigned long long n, unsigned long long n_plane, float4* __restrict__ pos,
                                 float4* __restrict__ attr) {
	unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float* v = raw + i * 14;
	pos[i] = make_float4(v[0], v[1], v[2], 1.0f);
	const float4 a0 = make_float4(v[3], v[4], v[5], v[6]), a1 = make_float4(v[7], v[8], v[9], v[10]), a2 = make_float4(v[11], v[12], v[13], 0.0f);
#if AXR_ATTR_PLANES
	attr[i] = a0; attr[n_plane + i] = a1; attr[2 * n_plane + i] = a2;
#else
	attr[3 * i] = a0; attr[3 * i + 1] = a1; attr[3 * i + 2] = a2;
#endif
}
// linear RGBA8 rows -> the tiled order the samplers read (axr_upload_texture); the padding texels are never addressed
__global__ void k_tile_texture(const uchar4* __restrict__ linear, int w, int h, int tiles_x, uchar4* __restrict__ tiled) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x < w && y < h) tiled[tex_offset_y(y, w, tiles_x) + tex_offset_x(x)] = linear[(size_t)y * w + x];
}
// 12 B index triples -> padded 16 B quads (AXR_IDX_PAD)
__global__ void k_pad_indices(const unsigned* __restrict__ idx, unsigned long long n_faces, uint4* __restrict__ idx4) {
	unsigned long long f = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (f < n_faces) idx4[f] = make_uint4(idx[3 * f], idx[3 * f + 1], idx[3 * f + 2], 0u);
}

// ------------------------------------------------------------------------------------------------ FP32 issue micro-benchmark
// the shading / setup kernels are ALU-heavy, so every report carries an FP32-issue bound next to the HBM one.
// Eight independent chains per thread; MODE 0: x = x*a + b as separate FMUL + FADD (what the path executes: the library is built
// with -fmad=false for parity), MODE 1: fused (__fmaf_rn). The result is stored so nothing is optimised away.
template <int MODE>
__global__ void __launch_bounds__(256) k_fp32_peak(float* out, float a, float b, int iters) {
	float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int k = 0; k < 16; ++k) {
			if (MODE == 0) {
				x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
				x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
			} else {
				x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b);
				x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b);
			}
		}
	}
	out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// ------------------------------------------------------------------------------------------------ random-gather micro-benchmark
// The shading stage reads its vertices and texels as data-dependent 16-byte gathers, one 32-byte DRAM sector each when nothing is
// shared; how many such sectors per second the memory system delivers is the bound that matters for it, not the streaming bandwidth.
// Every thread issues `ILP` independent 16-byte loads per step at pseudo-random 32-byte-aligned offsets of a buffer much larger
// than L2 (the offsets of a warp's lanes are unrelated: 32 sectors per warp-wide load), and chains the next step's offsets on the
// loaded data so that nothing can be hoisted.
// SPAN: consecutive sectors fetched per random position (1: a lone 32-byte sector; 2: a 64-byte record; 4: a whole 128-byte line)
template <int ILP, int SPAN = 1>
__global__ void __launch_bounds__(256) k_gather_peak(const uint4* __restrict__ buf, unsigned long long n_sectors, int steps, unsigned* out) {
	unsigned long long x = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
	unsigned acc = 0;
	for (int s = 0; s < steps; ++s) {
		uint4 v[ILP];
#pragma unroll
		for (int k = 0; k < ILP; ++k) {
			x = x * 6364136223846793005ull + 1442695040888963407ull;
			const uint4* p = buf + (((x >> 20) % n_sectors) / SPAN) * (2ull * SPAN);  // aligned to SPAN sectors
			v[k] = __ldg(p);  // the first half of a sector
#pragma unroll
			for (int j = 1; j < SPAN; ++j) { const uint4 w = __ldg(p + 2 * j); v[k].x ^= w.x; v[k].w ^= w.y; }
		}
#pragma unroll
		for (int k = 0; k < ILP; ++k) acc += v[k].x ^ v[k].w;
		x += acc & 1u;
	}
	out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ------------------------------------------------------------------------------------------------ vertex stage
// reference src/tiled_pipeline.cpp:210-212 (mvp * vec4(pos,1)) + :57-66 (perspective divide), once per unique vertex
// It also zeroes the draw's device counters (k_setup_raster, the next kernel on the stream, is their first user), which
// saves two memset nodes per draw.
__global__ void __launch_bounds__(256) k_vertex_xform(const float4* __restrict__ pos, unsigned long long n, m4 mvp, float fW,
                                                      float fH, float4* __restrict__ sv, DrawStatus* status, unsigned* n_records, unsigned* n_clip_tiles,
                                                      unsigned* n_clip_faces) {
	const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < sizeof(DrawStatus) / 4) reinterpret_cast<unsigned*>(status)[t] = 0u;
	if (t == 0) { *n_records = 0u; *n_clip_tiles = 0u; *n_clip_faces = 0u; }
	// AXR_VERTEX_PER_THREAD vertices per thread, a whole grid apart (coalesced), all loads in flight before the first transform
	const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
	float4 p[AXR_VERTEX_PER_THREAD];
#pragma unroll
	for (int k = 0; k < AXR_VERTEX_PER_THREAD; ++k) {
		const unsigned long long i = t + k * stride;
		if (i < n) p[k] = __ldg(pos + i);
	}
#pragma unroll
	for (int k = 0; k < AXR_VERTEX_PER_THREAD; ++k) {
		const unsigned long long i = t + k * stride;
		if (i >= n) continue;
		v4 c = mul(mvp, V4(p[k].x, p[k].y, p[k].z, 1.0f));
		float sx, sy, z;
		to_screen(c, fW, fH, sx, sy, z);
		// a "safely outside" bit implies the exact bit of the same plane, so the margins are only evaluated for the few vertices
		// that are outside some plane at all
		const unsigned code = clip_code(c);
		sv[i] = make_float4(sx, sy, z, __uint_as_float(code ? (code | (clip_code_safe_out(c) << 8)) : 0u));
	}
}

// ------------------------------------------------------------------------------------------------ setup + small raster
struct SetupOut {
	unsigned long long* vis;     // W*H visibility keys
	unsigned* tile_touched;      // per GPU tile: 1 when the direct path wrote a key into it
	unsigned* tile_count;        // per GPU tile: binned references
	TriRecord* records;
	unsigned rec_cap;
	unsigned* n_records;         // device counter (also the number wanted when it overflows)
	DrawStatus* status;
	int small_dim, small_area;   // direct-raster limits of this draw
	int bins_enabled;            // BINS_*: how triangles too large for the direct path reach their tiles in this draw
	const unsigned long long* floor;  // depth peeling only (PEEL): per pixel, keys <= floor have been dealt with
	unsigned* clip_faces;        // faces that need the clipper: appended by k_setup_raster, worked through by k_setup_clipped
	unsigned* n_clip_faces;      // (capacity: the face count of the mesh, so the list cannot overflow)
	unsigned n_chunks, swz_rows; // CTA -> face chunk interleave of k_setup_raster (setup_grid())
	int banded;                  // the context owns rows [band_lo, band_hi) of the frame only
	float band_lo, band_hi;
};
constexpr unsigned OVF_RECORDS = 1u, OVF_REFS = 2u, OVF_NEED_BINS = 4u;
// BINS_NONE:  the mesh had no such triangle last time: no records, no bin kernels; one that turns up is counted, k_setup_clipped raises
//             OVF_NEED_BINS and the host re-issues the draw with BINS_LISTS.
// BINS_LISTS: records + per-tile counts -> k_scan_tiles -> k_bin_scatter -> per-tile reference lists.
// BINS_SCAN:  the mesh had few of them last time (<= BINS_SCAN_MAX): records + per-tile counts only; a tile CTA with a non-zero count
//             tests every record against its tile itself, which is cheaper than two more kernels in the draw's dependency chain
//             (C1: 17 us of a 95 us draw). More records than BINS_SCAN_MAX (or than fit) -> OVF_NEED_BINS -> re-issued with lists.
enum { BINS_NONE = 0, BINS_LISTS = 1, BINS_SCAN = 2 };
constexpr unsigned BINS_SCAN_MAX = 4096;

struct EmitCounters { unsigned tris, small, binned; };

// Mark GPU tile (tx,ty) as holding direct-path keys. Test before set: millions of plain stores to the same few thousand
// flags serialise in L2 (measured: 86 us of a 246 us kernel on C3); the flag is monotonic within a draw, so a stale cached 0
// only costs a redundant store.
__device__ __forceinline__ void touch_tile(const FrameParams& fp, const SetupOut& o, int tx, int ty) {
	unsigned* tf = o.tile_touched + ty * fp.ntx + tx;
	if (*tf == 0u) *tf = 1u;
}

#ifndef AXR_SETUP_ABLATE
#define AXR_SETUP_ABLATE 0  // measurement only (tools/build_variants.py): 1 = loads + cull, 2 = + triangle setup, 3 = + coverage loop without the reductions
#endif
#ifndef AXR_SETUP_LOOP
#define AXR_SETUP_LOOP 1  // 1: runs x rows x pixels with the run / row terms hoisted (C3: 177 us; 2: four pixels per step, branch-free: 195 us); 0: closed-form coverage() per pixel in one counted loop
#endif

#ifndef AXR_SETUP_TRIM
#define AXR_SETUP_TRIM 0  // instruction trims of the pixel loop: bit 0 = float lane counter, bit 1 = one NaN-propagating 3-input minimum + one compare
#endif
// c0 >= 0 && c1 >= 0 && c2 >= 0 with NaN failing (_CMP_GE_OQ). min.NaN returns NaN when any input is one, and NaN >= 0 is false;
// otherwise the minimum is >= 0 exactly when all three are (-0 >= 0 holds, whichever zero the minimum picks).
__device__ __forceinline__ bool all_ge0(float c0, float c1, float c2) {
#if defined(__CUDA_ARCH__) && (AXR_SETUP_TRIM & 2)
	float m;
	asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(c0), "f"(c1), "f"(c2));
	return m >= 0.f;
#else
	return c0 >= 0.f && c1 >= 0.f && c2 >= 0.f;
#endif
}
// inverse of small_int_to_f32 for 0 <= f < 2^22, f an integer
__device__ __forceinline__ int f32_to_small_int(float f) {
#ifdef __CUDA_ARCH__
	return (int)(__float_as_uint(f + 8388608.0f) - 0x4B000000u);
#else
	return (int)f;
#endif
}

// Exact coverage + visibility keys of a triangle whose pixel box is small (at most 12 x 12, 64 px), by the thread that set it up.
// Returns true when a key was written.
template <bool PEEL>
__device__ __forceinline__ bool raster_small(const FrameParams& fp, const SetupOut& o, const Setup& s, unsigned ordinal) {
	bool any = false;
	auto hit = [&](int px, int py, float c0, float c1, float c2) {
		float al, be, ga;
		const float z = interp_z(s, c0, c1, c2, al, be, ga);
		if (!z_draws(z)) return;
		const unsigned long long key = make_key(z, ordinal);
		unsigned long long* slot = o.vis + ((unsigned)py * (unsigned)fp.W + (unsigned)px);
		if constexpr (PEEL) {
			if (!(key > o.floor[(unsigned)py * (unsigned)fp.W + (unsigned)px])) return;
		}
#if AXR_SETUP_ABLATE != 3
		atomicMin(slot, key);  // result unused -> RED.MIN.64, fire and forget
#endif
		any = true;
	};
#if AXR_SETUP_LOOP == 0
	// One counter over the whole pixel box with the closed-form coverage() per pixel
	const int bw = s.X1 - s.X0, bh = s.Y1 - s.Y0;
	int px = s.X0, py = s.Y0;
	for (int i = bw * bh; i > 0; --i) {
		float c0, c1, c2;
		if (coverage(s, px, py, c0, c1, c2)) hit(px, py, c0, c1, c2);
		if (++px == s.X1) { px = s.X0; ++py; }
	}
#else
	// coverage() (axr_raster.cuh) is the specification; this walks the same pixels with the terms that do not change hoisted, every
	// + and * being the one coverage() performs for that pixel. A row of the box splits into at most three RUNS of pixels that share
	// startX = max(px & ~15, fminx) and the `d >= 8` branch: a run ends at the next 16-px reference tile and, while d < 8, at
	// startX + 8. Per run: a*sxc; per row of a run: ((a*sxc + b*pyc) + c) [+ a*8]; per pixel: + a*(float)i and the three compares.
	int xs = s.X0;
#pragma unroll 1
	do {
		const int startX = max(xs & ~(REF_TILE - 1), s.fminx);
		const bool hi = xs - startX >= 8;  // the second group of eight of the reference's 16-px row
		int xe = min(s.X1, (xs | (REF_TILE - 1)) + 1);
		if (!hi) xe = min(xe, startX + 8);
		const float sxc = small_int_to_f32(startX) + 0.5f;
		const float p0 = s.a0 * sxc, p1 = s.a1 * sxc, p2 = s.a2 * sxc;
		const int i0 = hi ? startX + 8 : startX;  // the pixel whose lane index i is 0
#if AXR_SETUP_TRIM & 1
		const float fi0 = small_int_to_f32(xs - i0), fend = small_int_to_f32(xe - i0);
#endif
#pragma unroll 1
		for (int py = s.Y0; py < s.Y1; ++py) {
			const float pyc = small_int_to_f32(py) + 0.5f;
			float r0 = p0 + s.b0 * pyc + s.c0, r1 = p1 + s.b1 * pyc + s.c1, r2 = p2 + s.b2 * pyc + s.c2;
			// a * 8 is exact (a power of two), so the fused form rounds once, exactly where r + a*8 does: same bits, one instruction
			if (hi) { r0 = __fmaf_rn(s.a0, 8.0f, r0); r1 = __fmaf_rn(s.a1, 8.0f, r1); r2 = __fmaf_rn(s.a2, 8.0f, r2); }
#if AXR_SETUP_LOOP == 2
			// four pixels at a time, branch-free: twelve independent chains instead of one pixel's dependent one (the loop is bound by
			// the latency of its own chain, not by issue slots); the rare covered pixel recomputes its three values in hit4()
#pragma unroll 1
			for (int px0 = xs; px0 < xe; px0 += 4) {
				unsigned m = 0;
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					const float fi = small_int_to_f32(px0 - i0 + k);
					const float c0 = r0 + s.a0 * fi, c1 = r1 + s.a1 * fi, c2 = r2 + s.a2 * fi;
					if (c0 >= 0.f && c1 >= 0.f && c2 >= 0.f && px0 + k < xe) m |= 1u << k;
				}
				while (m) {
					const int px = px0 + (__ffs(m) - 1);
					m &= m - 1;
					const float fi = small_int_to_f32(px - i0);
					hit(px, py, r0 + s.a0 * fi, r1 + s.a1 * fi, r2 + s.a2 * fi);
				}
			}
#elif AXR_SETUP_TRIM & 1
			// the lane index itself is the loop counter (small integers are exact in binary32); the pixel column is only rebuilt for a hit
#pragma unroll 1
			for (float fi = fi0; fi < fend; fi += 1.0f) {
				const float c0 = r0 + s.a0 * fi, c1 = r1 + s.a1 * fi, c2 = r2 + s.a2 * fi;
				if (all_ge0(c0, c1, c2)) hit(i0 + f32_to_small_int(fi), py, c0, c1, c2);
			}
#else
#pragma unroll 1
			for (int px = xs; px < xe; ++px) {
				const float fi = small_int_to_f32(px - i0);
				const float c0 = r0 + s.a0 * fi, c1 = r1 + s.a1 * fi, c2 = r2 + s.a2 * fi;
				if (all_ge0(c0, c1, c2)) hit(px, py, c0, c1, c2);
			}
#endif
		}
		xs = xe;
	} while (xs < s.X1);
#endif
	return any;
}

// DEFER_TOUCH: the caller flags the tiles later (warp-aggregated); the return value is the tile rect of the pixel box packed as
// tx0 | ty0<<8 | tx1<<16 | ty1<<24 in units of GPU tiles (frames up to 8160 px), or NO_TOUCH when there is nothing to flag.
constexpr unsigned NO_TOUCH = 0xFFFFFFFFu;
template <bool PEEL, bool DEFER_TOUCH, bool IN_FRAME = false>
__device__ __forceinline__ unsigned emit_triangle(const FrameParams& fp, const SetupOut& o, float x0, float y0, float x1, float y1,
                                                  float x2, float y2, float z0, float z1, float z2, unsigned ordinal, EmitCounters& cnt) {
	cnt.tris++;
#if AXR_SETUP_ABLATE == 1
	return NO_TOUCH;
#endif
	Setup s;
	if (!setup_triangle<IN_FRAME>(x0, y0, x1, y1, x2, y2, z0, z1, z2, fp.W, fp.y_lo, fp.y_hi, s)) return NO_TOUCH;
	const int bw = s.X1 - s.X0, bh = s.Y1 - s.Y0;
	if (bw <= o.small_dim && bh <= o.small_dim && bw * bh <= o.small_area) {
		cnt.small++;
#if AXR_SETUP_ABLATE == 2
		if (s.inv_area == 123.0f) cnt.small++;
		return NO_TOUCH;
#endif
		if (raster_small<PEEL>(fp, o, s, ordinal)) {
			const int tx0 = s.X0 / GT, ty0 = s.Y0 / GT, tx1 = (s.X1 - 1) / GT, ty1 = (s.Y1 - 1) / GT;  // box <= 12x12 px: at most 2x2 tiles
			if (DEFER_TOUCH && fp.ntx < 256 && fp.nty < 256)  // strictly: tile (255,255) alone would pack to NO_TOUCH
				return (unsigned)tx0 | ((unsigned)ty0 << 8) | ((unsigned)tx1 << 16) | ((unsigned)ty1 << 24);
			for (int ty = ty0; ty <= ty1; ++ty)
				for (int tx = tx0; tx <= tx1; ++tx) touch_tile(fp, o, tx, ty);
		}
		return NO_TOUCH;
	}
	cnt.binned++;
	if (!o.bins_enabled) return NO_TOUCH;  // counted: k_setup_clipped raises OVF_NEED_BINS and the draw is re-issued with the bin kernels
	// warp-aggregated append of the setup record
	cg::coalesced_group g = cg::coalesced_threads();
	unsigned base = 0;
	if (g.thread_rank() == 0) base = atomicAdd(o.n_records, g.size());
	unsigned slot = g.shfl(base, 0) + g.thread_rank();
	if (slot < o.rec_cap) {
		TriRecord r;
		r.x0 = x0; r.y0 = y0; r.x1 = x1; r.y1 = y1; r.x2 = x2; r.y2 = y2; r.z0 = z0; r.z1 = z1; r.z2 = z2; r.ordinal = ordinal;
		o.records[slot] = r;
	}
	int tx0 = s.X0 / GT, tx1 = (s.X1 - 1) / GT, ty0 = s.Y0 / GT, ty1 = (s.Y1 - 1) / GT;
	for (int ty = ty0; ty <= ty1; ++ty)
		for (int tx = tx0; tx <= tx1; ++tx) atomicAdd(o.tile_count + ty * fp.ntx + tx, 1u);
	return NO_TOUCH;
}

// Clip slow path of one face for the visibility pass (positions only): reference src/tiled_pipeline.cpp:210-234.
// Returns its counters packed as tris | small << 8 | binned << 16 (at most 8 sub-triangles per face): handing the caller's
// counters to this out-of-line function by reference would pin them in local memory on the hot path as well.
template <bool PEEL>
__device__ __noinline__ unsigned setup_clipped_face(const FrameParams& fp, const SetupOut& o, const m4& mvp, float4 p0, float4 p1,
                                                    float4 p2, unsigned face) {
	EmitCounters cnt = {0, 0, 0};
	ClipPos a[MAX_CLIPPED_VERTS], b[MAX_CLIPPED_VERTS];
	a[0].clip = mul(mvp, V4(p0.x, p0.y, p0.z, 1.0f));
	a[1].clip = mul(mvp, V4(p1.x, p1.y, p1.z, 1.0f));
	a[2].clip = mul(mvp, V4(p2.x, p2.y, p2.z, 1.0f));
	ClipPos* out;
	int n = clip_triangle(a, b, &out, clip_planes_needed(a[0].clip, a[1].clip, a[2].clip));
	const float fW = (float)fp.W, fH = (float)fp.H;
	for (int j = 0; j + 2 < n; j += 3) {
		float x0, y0, z0, x1, y1, z1, x2, y2, z2;
		to_screen(out[j].clip, fW, fH, x0, y0, z0);
		to_screen(out[j + 1].clip, fW, fH, x1, y1, z1);
		to_screen(out[j + 2].clip, fW, fH, x2, y2, z2);
		if (is_backface(x0, y0, x1, y1, x2, y2)) continue;
		emit_triangle<PEEL, false>(fp, o, x0, y0, x1, y1, x2, y2, z0, z1, z2, face * 8u + (unsigned)(j / 3), cnt);
	}
	return cnt.tris | (cnt.small << 8) | (cnt.binned << 16);
}

#ifndef AXR_SETUP_THREADS
#define AXR_SETUP_THREADS 128
#endif
constexpr int SETUP_THREADS = AXR_SETUP_THREADS;
#ifndef AXR_SETUP_MINB
#define AXR_SETUP_MINB 12  // 40 registers, 48 resident warps (16 x 32 registers spills in the coverage loop: C3 197 against 175 us)
#endif
#ifndef AXR_TILE_MINB
#define AXR_TILE_MINB 4
#endif
#ifndef AXR_QUAD_LANES
#define AXR_QUAD_LANES 1  // C3: 158 -> 154 us
#endif
#ifndef AXR_TILE_IDX_STASH
#define AXR_TILE_IDX_STASH 0
#endif
#ifndef AXR_TILE_RECOMPUTE_SV
#define AXR_TILE_RECOMPUTE_SV 0
#endif

// The draw's counters go to the host through mapped pinned memory (plain stores over PCIe): a cudaMemcpyAsync between the
// kernels would put a copy-engine operation, i.e. a bubble of ~10 us, into the middle of every draw.
// No system fence: the host reads after synchronising on an event recorded behind the publishing kernel.
__device__ __forceinline__ void publish_status(const DrawStatus* d, DrawStatus* h, int word) {
	reinterpret_cast<volatile unsigned long long*>(h)[word] = reinterpret_cast<const volatile unsigned long long*>(d)[word];
}

constexpr int SETUP_CHUNK = SETUP_THREADS;  // faces per CTA, one per thread
#ifndef AXR_SETUP_SWZ_K
#define AXR_SETUP_SWZ_K 16
#endif
#ifndef AXR_SETUP_SWZ_GROUP
#define AXR_SETUP_SWZ_GROUP 128
#endif
#ifndef AXR_SETUP_PF
#define AXR_SETUP_PF 1776  // k_setup_raster: L2 prefetch of the index chunk of the CTA this many positions ahead (0: off); 888 / 3552: the same
#endif
constexpr unsigned SETUP_SWZ_K = AXR_SETUP_SWZ_K, SETUP_SWZ_GROUP = AXR_SETUP_SWZ_GROUP;
// grid of k_setup_raster for n_faces faces; fills the interleave parameters
inline unsigned setup_grid(unsigned long long n_faces, unsigned& n_chunks, unsigned& swz_rows) {
	n_chunks = (unsigned)((n_faces + SETUP_CHUNK - 1) / SETUP_CHUNK);
	swz_rows = 0;
	if (SETUP_SWZ_K <= 1) return n_chunks;
	const unsigned groups = (n_chunks + SETUP_SWZ_GROUP - 1) / SETUP_SWZ_GROUP;
	swz_rows = (groups + SETUP_SWZ_K - 1) / SETUP_SWZ_K;
	return swz_rows * SETUP_SWZ_K * SETUP_SWZ_GROUP;
}

// One face per thread. (Measured and rejected: two or four faces per thread with the later faces' indices staged in shared memory
// and their screen records prefetched into L1 — the load phase alone gets faster, 70 -> 58 us on C3, the kernel does not: 166-194
// against 161 us; the clipper called from here instead of from its own kernel: 178 us.)
template <bool PEEL>
__global__ void __launch_bounds__(SETUP_THREADS, AXR_SETUP_MINB) k_setup_raster(const __grid_constant__ MeshView mesh, const float4* __restrict__ sv,
                                                                const __grid_constant__ FrameParams fp, const __grid_constant__ SetupOut o) {
	const unsigned lane = threadIdx.x & 31u;
	// CTA -> chunk: the hardware hands out CTAs in index order, so the ~1800 resident ones would all sit in one stretch of the index
	// buffer, and meshes are laid out coherently: whole stretches are back-facing (their warps only wait for loads) or front-facing
	// (their warps only compute), the two phases alternate GPU-wide and never overlap. Groups of SETUP_SWZ_GROUP consecutive chunks are
	// therefore dealt out round-robin over SETUP_SWZ_K distant regions of the mesh.
	unsigned chunk = blockIdx.x;
	if (SETUP_SWZ_K > 1) {
		const unsigned g = blockIdx.x / SETUP_SWZ_GROUP, j = blockIdx.x % SETUP_SWZ_GROUP;
		chunk = ((g % SETUP_SWZ_K) * o.swz_rows + g / SETUP_SWZ_K) * SETUP_SWZ_GROUP + j;
		if (chunk >= o.n_chunks) return;
	}
#if defined(__CUDA_ARCH__) && AXR_SETUP_PF > 0
	// Software pipeline across CTAs (the hardware starts them in index order, 148 x 12 are resident): the index triples of the chunk
	// that the CTA one generation further on will read (12 lines, one thread each) are asked into L2 now, so that CTA's first round trip
	// is an L2 hit. C3 155 -> 152-154 us, C4 249 -> 245 us. (Measured and rejected: a second stage that reads those indices one
	// generation early and asks for the screen records they name, 162 us: three more loads and three prefetches per thread cost more
	// request slots than the latency they hide; the same idea for k_tile_shade's keys, 137 -> 141 us.)
	{
		const unsigned fb = blockIdx.x + AXR_SETUP_PF;
		unsigned fc = fb;
		if (SETUP_SWZ_K > 1) {
			const unsigned g = fb / SETUP_SWZ_GROUP, j = fb % SETUP_SWZ_GROUP;
			fc = ((g % SETUP_SWZ_K) * o.swz_rows + g / SETUP_SWZ_K) * SETUP_SWZ_GROUP + j;
		}
		if (threadIdx.x < SETUP_CHUNK * 12 / 128 && fb < gridDim.x && fc < o.n_chunks && (fc + 1u) * SETUP_CHUNK <= (unsigned)mesh.n_faces)
			asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(mesh.idx + 3u * fc * SETUP_CHUNK) + threadIdx.x * 128u));
	}
#endif
	const unsigned f = chunk * SETUP_CHUNK + threadIdx.x;  // faces < 2^29 (axr_upload_mesh): 32-bit index arithmetic throughout
	EmitCounters cnt = {0, 0, 0};  // each 0 or 1 here: a face that is not clipped is one triangle
	unsigned touched = NO_TOUCH;   // tile rect the direct path wrote keys into (packed, see emit_triangle)
	bool clip = false;
	if (f < (unsigned)mesh.n_faces && f >= mesh.first_face) {
		const unsigned* ip = mesh.idx + 3u * f;
		const unsigned i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2);
		const float4 s0 = __ldg(sv + i0), s1 = __ldg(sv + i1), s2 = __ldg(sv + i2);
		const unsigned k0 = __float_as_uint(s0.w), k1 = __float_as_uint(s1.w), k2 = __float_as_uint(s2.w);
		if (((k0 | k1 | k2) & 0x3fu) == 0) {
			// every vertex inside every plane: clipTriangle returns the triangle unchanged (reference src/pipeline.cpp:322-325)
			// A context that owns a band of the frame (multi-GPU screen bands) first drops the faces whose rows miss it: the pixel box
			// [max(y_lo, floor(minY)), min(y_hi, ceil(maxY))) is empty exactly when maxY <= y_lo or minY >= y_hi (y_lo, y_hi integers), so
			// this is setup_triangle()'s own early return, taken before the back-face test and the box arithmetic for the 1 - 1/N of
			// the replicated geometry that belongs to other ranks. (Such faces are not counted in the draw's statistics.)
			const bool in_band = !o.banded || !(max3f(s0.y, s1.y, s2.y) <= o.band_lo || min3f(s0.y, s1.y, s2.y) >= o.band_hi);
			if (in_band && !is_backface(s0.x, s0.y, s1.x, s1.y, s2.x, s2.y))
				touched = emit_triangle<PEEL, true, true>(fp, o, s0.x, s0.y, s1.x, s1.y, s2.x, s2.y, s0.z, s1.z, s2.z, f * 8u, cnt);
		} else if ((k0 & k1 & k2) >> 8) {
			// all three vertices safely outside one plane: clipTriangle returns nothing (see clip_code_safe_out)
		} else {
			clip = true;  // needs the clipper: k_setup_clipped, the next kernel on the stream
		}
	}
	__syncwarp();
	// Faces for the clipper: warp-aggregated append (any order: keys carry the face ordinal, bins are order-free)
	const unsigned cm = __ballot_sync(0xffffffffu, clip);
	if (cm) {
		unsigned at = 0;
		if (lane == 0) at = atomicAdd(o.n_clip_faces, (unsigned)__popc(cm));
		at = __shfl_sync(0xffffffffu, at, 0);
		if (clip) o.clip_faces[at + __popc(cm & ((1u << lane) - 1u))] = f;
	}
	// Tile flags of the direct path: the 32 consecutive faces of a warp mostly land in the same tile (or the same pair), and then one
	// lane flags it for all of them; otherwise every lane flags its own (test before set, see touch_tile).
	const unsigned hm = __ballot_sync(0xffffffffu, touched != NO_TOUCH);
	if (hm) {
		const unsigned first = __shfl_sync(0xffffffffu, touched, __ffs(hm) - 1);
		const bool uniform = __ballot_sync(0xffffffffu, touched == NO_TOUCH || touched == first) == 0xffffffffu;
		if (uniform ? lane == 0 : touched != NO_TOUCH) {
			const unsigned r = uniform ? first : touched;
			const int tx1 = (int)((r >> 16) & 255u), ty1 = (int)(r >> 24);
			for (int ty = (int)((r >> 8) & 255u); ty <= ty1; ++ty)
				for (int tx = (int)(r & 255u); tx <= tx1; ++tx) touch_tile(fp, o, tx, ty);
		}
	}
	// Counters: one warp sum of the three packed counters (each at most 32), then fire-and-forget reductions into one of STAT_STRIPES
	// stripes (no hot address): triangles and small triangles sit next to each other and go up with one 64-bit reduction. Warps whose
	// faces were all culled have nothing to add.
	if (__ballot_sync(0xffffffffu, cnt.tris != 0u)) {
		const unsigned w = __reduce_add_sync(0xffffffffu, cnt.tris | (cnt.small << 8) | (cnt.binned << 16));
		if (lane == 0) {
			unsigned* st = o.status->stripes[(chunk * (SETUP_THREADS / 32) + (threadIdx.x >> 5)) % STAT_STRIPES];
			atomicAdd(reinterpret_cast<unsigned long long*>(st + STRIPE_TRIS), (unsigned long long)(w & 255u) | ((unsigned long long)((w >> 8) & 255u) << 32));
			if (w >> 16) atomicAdd(st + STRIPE_BINNED, w >> 16);
		}
	}
}

// The faces k_setup_raster left for the clipper (usually a few thousand, often none), one thread per face; the CTA that finishes
// last then folds the striped counters of both kernels and publishes the draw's status to the host. (The ticket lives in this small
// kernel: inside k_setup_raster its barrier keeps every CTA resident until its slowest warp is done, C3 171 -> 239 us.)
constexpr int CLIPSETUP_THREADS = 64, CLIPSETUP_CTAS = 148;
template <bool PEEL>
__global__ void __launch_bounds__(CLIPSETUP_THREADS) k_setup_clipped(const __grid_constant__ MeshView mesh, const __grid_constant__ m4 mvp,
                                                                    const __grid_constant__ FrameParams fp, const __grid_constant__ SetupOut o,
                                                                    DrawStatus* host_status) {
	__shared__ unsigned long long s_fold[4][CLIPSETUP_THREADS / 32];
	__shared__ unsigned s_last;
	const unsigned n = *o.n_clip_faces;
	unsigned tris = 0, small = 0, binned = 0;
	for (unsigned i = blockIdx.x * CLIPSETUP_THREADS + threadIdx.x; i < n; i += gridDim.x * CLIPSETUP_THREADS) {
		const unsigned f = o.clip_faces[i];
		const unsigned* ip = mesh.idx + 3u * f;
		const unsigned c = setup_clipped_face<PEEL>(fp, o, mvp, __ldg(mesh.pos + __ldg(ip)), __ldg(mesh.pos + __ldg(ip + 1)), __ldg(mesh.pos + __ldg(ip + 2)), f);
		tris += c & 255u; small += (c >> 8) & 255u; binned += c >> 16;
	}
	if (tris) {
		unsigned* st = o.status->stripes[(blockIdx.x * CLIPSETUP_THREADS + threadIdx.x) % STAT_STRIPES];
		atomicAdd(st + STRIPE_TRIS, tris);
		if (small) atomicAdd(st + STRIPE_SMALL, small);
		if (binned) atomicAdd(st + STRIPE_BINNED, binned);
	}
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) s_last = atomicAdd(&o.status->pad, 1u) == gridDim.x - 1 ? 1u : 0u;
	__syncthreads();
	if (!s_last) return;
	__threadfence();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
	for (int c = 0; c < 3; ++c) {
		unsigned long long acc = 0;
		for (int i = threadIdx.x; i < STAT_STRIPES; i += CLIPSETUP_THREADS) acc += __ldcg(&o.status->stripes[i][c]);
		for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
		if (lane == 0) s_fold[c][warp] = acc;
	}
	__syncthreads();
	if (threadIdx.x < 4) {
		// status words: clipped_faces (= the list), triangles, small_triangles, binned_triangles
		unsigned long long sum = n;
		if (threadIdx.x) {
			sum = 0;
			for (int w = 0; w < CLIPSETUP_THREADS / 32; ++w) sum += s_fold[threadIdx.x - 1][w];
		}
		(&o.status->clipped_faces)[threadIdx.x] = sum;
		if (threadIdx.x == 3) {
			if (o.bins_enabled == BINS_NONE && sum != 0) o.status->overflow = OVF_NEED_BINS;
			if (o.bins_enabled == BINS_SCAN && (sum > BINS_SCAN_MAX || sum > o.rec_cap)) o.status->overflow = OVF_NEED_BINS;
		}
	}
	__syncthreads();
	// bin_refs and (with bins) overflow are k_scan_tiles' to fill in; it publishes those words again
	if (threadIdx.x < 6) publish_status(o.status, host_status, threadIdx.x);
}

// ------------------------------------------------------------------------------------------------ bins: scan + scatter
// Block-wide exclusive prefix sum over the per-tile counts (one CTA of 1024 threads, 8 tiles per thread per round).
// On exit: bin_start[0..n] holds offsets, tile_count[] is zeroed so k_bin_scatter can reuse it as the fill cursor.
// Both bin kernels are only launched for draws issued with bins (SetupOut::bins_enabled); with no record at all they return at once.
constexpr int SCAN_THREADS = 1024, SCAN_PER_THREAD = 8;
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(unsigned* tile_count, unsigned* bin_start, int n, unsigned ref_cap,
                                                             const unsigned* n_records, unsigned rec_cap, DrawStatus* status, DrawStatus* host_status) {
	__shared__ unsigned s_warp[32];
	__shared__ unsigned s_carry;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const unsigned nrec = *n_records;
	if (nrec == 0) return;  // k_tile_shade does not read bin_start in this case; the status was published by k_setup_raster
	if (tid == 0) s_carry = 0;
	__syncthreads();
	for (int base = 0; base < n; base += SCAN_THREADS * SCAN_PER_THREAD) {
		const int i0 = base + tid * SCAN_PER_THREAD;
		unsigned v[SCAN_PER_THREAD], sum = 0;
#pragma unroll
		for (int k = 0; k < SCAN_PER_THREAD; ++k) { v[k] = (i0 + k < n) ? tile_count[i0 + k] : 0u; sum += v[k]; }
		unsigned x = sum;
		for (int d = 1; d < 32; d <<= 1) {
			unsigned y = __shfl_up_sync(0xffffffffu, x, d);
			if (lane >= d) x += y;
		}
		if (lane == 31) s_warp[warp] = x;
		__syncthreads();
		if (warp == 0) {
			unsigned w = s_warp[lane];
			for (int d = 1; d < 32; d <<= 1) {
				unsigned y = __shfl_up_sync(0xffffffffu, w, d);
				if (lane >= d) w += y;
			}
			s_warp[lane] = w;
		}
		__syncthreads();
		const unsigned carry = s_carry;
		unsigned excl = carry + (warp ? s_warp[warp - 1] : 0u) + (x - sum);
#pragma unroll
		for (int k = 0; k < SCAN_PER_THREAD; ++k)
			if (i0 + k < n) { bin_start[i0 + k] = excl; tile_count[i0 + k] = 0u; excl += v[k]; }
		__syncthreads();
		if (tid == SCAN_THREADS - 1) s_carry = carry + s_warp[31];
		__syncthreads();
	}
	if (tid == 0) {
		unsigned total = s_carry;
		bin_start[n] = total;
		status->bin_refs = total;
		unsigned ovf = 0;
		if (nrec > rec_cap) ovf |= OVF_RECORDS;
		if (total > ref_cap) ovf |= OVF_REFS;
		status->overflow = ovf;
		__threadfence();
		publish_status(status, host_status, 4);  // bin_refs
		publish_status(status, host_status, 5);  // overflow
	}
}

__global__ void __launch_bounds__(256) k_bin_scatter(const TriRecord* __restrict__ records, const unsigned* __restrict__ n_records,
                                                     FrameParams fp, const unsigned* __restrict__ bin_start, unsigned* cursor,
                                                     unsigned* __restrict__ items, const DrawStatus* status) {
	if (status->overflow) return;
	unsigned n = *n_records;
	for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
		TriRecord t = records[r];
		Setup s;
		if (!setup_triangle(t.x0, t.y0, t.x1, t.y1, t.x2, t.y2, t.z0, t.z1, t.z2, fp.W, fp.y_lo, fp.y_hi, s)) continue;
		int tx0 = s.X0 / GT, tx1 = (s.X1 - 1) / GT, ty0 = s.Y0 / GT, ty1 = (s.Y1 - 1) / GT;
		for (int ty = ty0; ty <= ty1; ++ty)
			for (int tx = tx0; tx <= tx1; ++tx) {
				int tile = ty * fp.ntx + tx;
				unsigned slot = atomicAdd(cursor + tile, 1u);
				items[bin_start[tile] + slot] = r;
			}
	}
}

// ------------------------------------------------------------------------------------------------ per-tile raster + shade + resolve
struct TileIn {
	unsigned long long* vis;
	unsigned* tile_touched;
	unsigned* tile_cursor;          // == tile_count, reset here for the next draw
	const unsigned* bin_start;
	const unsigned* items;
	const TriRecord* records;
	const unsigned* n_records;
	const DrawStatus* status;
	const float4* sv;
	unsigned* color;                // BGRA8 as packed words (B | G<<8 | R<<16 | A<<24), full-frame pitch W
	float* depth;
	const float* depth_read;        // where the merge test reads fbZ from (== depth unless the output is write-only host memory)
	int read_depth;                 // 0: the caller guarantees depth == +inf everywhere (freshly cleared single-draw target)
	unsigned long long* floor;      // depth peeling only (Shader::DISCARDS): per-pixel key of the last discarded winner
	unsigned* again;                // depth peeling only: set when some winner was discarded in this pass
	int row_major;                  // 1: a warp shades one 32 x 1 pixel row (128 B contiguous stores: output in host memory over PCIe)
	unsigned* clip_tiles;           // tiles holding pixels owned by a clipped face: shaded by k_shade_clipped, not here
	unsigned* dirty;                // optional (axr_set_dirty_map): per GPU tile, set to 1 when this draw may store into the tile
	unsigned* n_clip_tiles;
	int bin_mode;                   // BINS_* of this draw
	// axr_set_output_fill: the draw overwrites EVERY pixel of the tiles it touches — the shaded colour where a triangle is visible,
	// (fill_color, fill_depth) elsewhere — and its merge test sees fill_depth: for a target that holds exactly one draw on top of a
	// clear (a multi-GPU composite slot), so that the target's owner never clears the tiles the next frame touches again
	int fill;
	unsigned fill_color;
	float fill_depth;
};

// (uint8)(int)(clamp(c, 0, 1) * 255): reference src/tiled_pipeline.cpp:579-582. cvttss2si turns NaN into 0x80000000, whose low byte is 0.
__device__ __forceinline__ unsigned to_u8(float c) {
	float v = clampf(c, 0.0f, 1.0f) * 255.0f;  // in [0, 255], or NaN
#ifdef __CUDA_ARCH__
	v = (v == v) ? v : 0.0f;
	return __float_as_uint(__fadd_rz(v, 8388608.0f)) & 0xffu;  // truncation without F2I (see floor_small)
#else
	return (unsigned)(unsigned char)cvtt(v);
#endif
}
__device__ __forceinline__ unsigned pack_bgra(v4 c) {
	// + the R<->B swizzle of mergeTileResults :1171-1174
	return to_u8(c.z) | (to_u8(c.y) << 8) | (to_u8(c.x) << 16) | (to_u8(c.w) << 24);
}

// Material group of a face: groups are ascending face ranges (reference src/mesh.cpp:336-346), group_first has n_groups + 1 entries
__device__ __forceinline__ const Material& face_material(const MeshView& mesh, unsigned face) {
	if (mesh.n_groups <= 1) return mesh.material0;
	int lo = 0;
	{
		int hi = mesh.n_groups;
		while (hi - lo > 1) {
			const int mid = (lo + hi) >> 1;
			if ((unsigned long long)face >= mesh.group_first[mid]) lo = mid; else hi = mid;
		}
	}
	return mesh.materials[lo];
}

// Accumulate one vertex's VertexOutput into the interpolated varyings: step k of ((v0*al) + (v1*be)) + v2*ga.
template <typename Shader>
__device__ __forceinline__ void accumulate_vertex(const Uniforms& u, int k, float w, const VIn& v, float* var) {
	float o[Shader::NV];
	Shader::vertex(u, v.pos, v.n, v.t, v.b, v.u, v.v, o);
#pragma unroll
	for (int i = 0; i < Shader::NV; ++i) var[i] = (k == 0) ? o[i] * w : var[i] + o[i] * w;
}

// IShader::vertex x3 + IShader::fragment for one pixel; returns fragment's discard flag (reference src/tiled_pipeline.cpp:571-577:
// a discarded fragment leaves depth and colour). FAST: the functor's fused form (colour within 1 LSB instead of bit-equal).
template <typename Shader, int SMP, bool FAST>
__device__ __forceinline__ bool run_shader(const MeshView& mesh, const Uniforms& u, const TileIn& in, unsigned face, size_t gi, float z,
                                           const float w[3], const VIn v[3]) {
	v4 col;
	const Material& mat = face_material(mesh, face);
	if constexpr (FAST && Shader::HAS_FAST) {
		if (Shader::template shade_fast<SMP>(u, mat, w, v, col)) return true;
	} else {
		float var[Shader::NV];
		accumulate_vertex<Shader>(u, 0, w[0], v[0], var);
		accumulate_vertex<Shader>(u, 1, w[1], v[1], var);
		accumulate_vertex<Shader>(u, 2, w[2], v[2], var);
		if (Shader::template fragment<SMP>(u, mat, var, col)) return true;
	}
	in.depth[gi] = z;
	in.color[gi] = pack_bgra(col);
	return false;
}

__device__ __forceinline__ VIn load_vertex(const MeshView& mesh, unsigned i) {
	const float4 p = __ldg(mesh.pos + i);
#if AXR_ATTR_PLANES
	const float4 a0 = __ldg(mesh.attr + i), a1 = __ldg(mesh.attr + mesh.n_plane + i), a2 = __ldg(mesh.attr + 2 * mesh.n_plane + i);
#else
	const float4* ap = mesh.attr + 3ull * i;
	const float4 a0 = __ldg(ap), a1 = __ldg(ap + 1), a2 = __ldg(ap + 2);
#endif
	VIn v;
	v.pos = V3(p.x, p.y, p.z);
	v.u = a0.x; v.v = a0.y;
	v.n = V3(a0.z, a0.w, a1.x);
	v.t = V3(a1.y, a1.z, a1.w);
	v.b = V3(a2.x, a2.y, a2.z);
	return v;
}

// Pixel whose visible triangle comes from a clipped face: re-derive sub-triangle (ordinal & 7) with full attributes
// (reference src/pipeline.cpp:176-272). Rare; runs in its own kernel (k_shade_clipped) so that its 3.5 KB of clip buffers and
// its call frame stay out of the tile kernel. Always the exact colour arithmetic.
template <typename Shader, int SMP>
__device__ bool shade_pixel_clipped(const MeshView& mesh, const Uniforms& u, const FrameParams& fp, const TileIn& in, unsigned ordinal,
                                    int px, int py) {
	ClipFull a[MAX_CLIPPED_VERTS], b[MAX_CLIPPED_VERTS];
	const unsigned* ip = mesh.idx + (size_t)(ordinal >> 3) * 3;
	for (int k = 0; k < 3; ++k) {
		const VIn v = load_vertex(mesh, __ldg(ip + k));
		a[k].pos = v.pos;
		a[k].clip = mul(u.mvp, V4(v.pos.x, v.pos.y, v.pos.z, 1.0f));
		a[k].uv[0] = v.u; a[k].uv[1] = v.v;
		a[k].n = v.n; a[k].t = v.t; a[k].b = v.b;
	}
	ClipFull* out;
	const int n = clip_triangle(a, b, &out, clip_planes_needed(a[0].clip, a[1].clip, a[2].clip));
	const int sub = (int)(ordinal & 7u);
	if (sub * 3 + 2 >= n) return false;
	const ClipFull* c = out + sub * 3;
	float sx[3], sy[3], sz[3];
	for (int k = 0; k < 3; ++k) to_screen(c[k].clip, (float)fp.W, (float)fp.H, sx[k], sy[k], sz[k]);
	Setup s;
	if (!setup_triangle(sx[0], sy[0], sx[1], sy[1], sx[2], sy[2], sz[0], sz[1], sz[2], fp.W, fp.y_lo, fp.y_hi, s)) return false;
	float c0, c1, c2, w[3];
	coverage(s, px, py, c0, c1, c2);
	const float z = interp_z(s, c0, c1, c2, w[0], w[1], w[2]);
	const size_t gi = (size_t)py * fp.W + px;
	if (!(z < (in.fill ? in.fill_depth : (in.read_depth ? in.depth_read[gi] : INFINITY)))) return false;
	VIn v[3];
	for (int k = 0; k < 3; ++k) { v[k].pos = c[k].pos; v[k].n = c[k].n; v[k].t = c[k].t; v[k].b = c[k].b; v[k].u = c[k].uv[0]; v[k].v = c[k].uv[1]; }
	return run_shader<Shader, SMP, false>(mesh, u, in, ordinal >> 3, gi, z, w, v);
}

// One visible pixel: all gathers that depend only on the vertex indices are issued together (screen records, positions,
// attributes, framebuffer depth), then edge setup -> barycentrics -> depth test -> shader.
enum { PIX_DONE = 0, PIX_DISCARDED = 1, PIX_CLIPPED = 2, PIX_LOST = 3 };  // stored / discarded by the shader / left to k_shade_clipped / lost the merge test
template <typename Shader, int SMP, bool FAST>
__device__ __forceinline__ int shade_pixel(const MeshView& mesh, const Uniforms& u, const FrameParams& fp, const TileIn& in, unsigned ordinal,
                                           unsigned i0, unsigned i1, unsigned i2, int px, int py, float fbz) {
	const size_t gi = (size_t)py * fp.W + px;
#if !AXR_TILE_RECOMPUTE_SV
	const float4 s0 = __ldg(in.sv + i0), s1 = __ldg(in.sv + i1), s2 = __ldg(in.sv + i2);
#endif
	VIn v[3];
	v[0] = load_vertex(mesh, i0); v[1] = load_vertex(mesh, i1); v[2] = load_vertex(mesh, i2);
#if AXR_TILE_RECOMPUTE_SV
	// variant: the vertex stage's 16 B screen records are recomputed from the positions (same arithmetic, same bits) instead of
	// gathered: three gathers fewer for ~150 instructions more
	float4 s0, s1, s2;
	{
		const float fW = (float)fp.W, fH = (float)fp.H;
		const v4 c0 = mul(u.mvp, V4(v[0].pos.x, v[0].pos.y, v[0].pos.z, 1.0f)), c1 = mul(u.mvp, V4(v[1].pos.x, v[1].pos.y, v[1].pos.z, 1.0f)),
		         c2 = mul(u.mvp, V4(v[2].pos.x, v[2].pos.y, v[2].pos.z, 1.0f));
		to_screen(c0, fW, fH, s0.x, s0.y, s0.z); to_screen(c1, fW, fH, s1.x, s1.y, s1.z); to_screen(c2, fW, fH, s2.x, s2.y, s2.z);
		s0.w = __uint_as_float(clip_code(c0)); s1.w = __uint_as_float(clip_code(c1)); s2.w = __uint_as_float(clip_code(c2));
	}
#endif
	if ((__float_as_uint(s0.w) | __float_as_uint(s1.w) | __float_as_uint(s2.w)) & 0x3fu) return PIX_CLIPPED;
	// the key exists, so the setup kernel's setup_triangle() succeeded for these very inputs: only the edge part is redone
	Setup s;
	s.fminx = cvtt(floorf(min3f(s0.x, s1.x, s2.x)));
	setup_edges(s0.x, s0.y, s1.x, s1.y, s2.x, s2.y, s0.z, s1.z, s2.z, s);
	float c0, c1, c2, w[3];
	coverage(s, px, py, c0, c1, c2);
	const float z = interp_z(s, c0, c1, c2, w[0], w[1], w[2]);
	// mergeTileResults: strict tileZ < fbZ (reference src/tiled_pipeline.cpp:1148-1156)
	if (!(z < fbz)) return PIX_LOST;
	return run_shader<Shader, SMP, FAST>(mesh, u, in, ordinal >> 3, gi, z, w, v) ? PIX_DISCARDED : PIX_DONE;
}

// candidates per round of the binned phase: 4 KB (+ 2 KB of candidate ids) of shared memory, so that keys + candidates of four resident CTAs stay inside the
// 64 KB shared-memory carve-out the keys alone already need (the gathers of the shading phase live on the rest of the L1)
constexpr int BIN_ROUND = 64, SCAN_CHUNK = 1024;
static_assert(BIN_ROUND <= TILE_THREADS && SCAN_CHUNK % TILE_THREADS == 0 && SCAN_CHUNK <= 65536, "one candidate per thread; 16-bit chunk-relative ids");
// FILL: axr_set_output_fill (TileIn::fill) as a compile-time switch, so that the plain kernel does not carry it (as a run-time flag it
// cost the C3 shading stage 6 us in registers and spills).
template <typename Shader, int SMP, bool FAST, bool FILL = false>
__global__ void __launch_bounds__(TILE_THREADS, AXR_TILE_MINB) k_tile_shade(const __grid_constant__ MeshView mesh, const __grid_constant__ Uniforms u, const __grid_constant__ FrameParams fp,
                                                                             const __grid_constant__ TileIn in) {
	constexpr bool PEEL = Shader::DISCARDS;
	__shared__ unsigned long long s_keys[GT_PIX];
	__shared__ float s_rec[BIN_ROUND][16];  // set-up candidates of one round (edges, 1/area, z, floor(minX), ordinal, tile-relative box)
	__shared__ unsigned short s_cand[SCAN_CHUNK];  // BINS_SCAN: records of the current chunk that reach into this tile
	__shared__ unsigned s_clipped, s_nrec, s_ncand;
	const int tx = blockIdx.x, ty = fp.ty_lo + blockIdx.y;
	const int tile = ty * fp.ntx + tx;
	const int x0 = tx * GT, y0 = ty * GT;
	// (the three loads that decide whether this CTA has anything to do are independent: issued together, one round trip for the
	// thousands of CTAs of an untouched tile)
	const unsigned ovf = in.status->overflow;
	const unsigned touched = in.tile_touched[tile];
	const unsigned nrec = in.bin_mode == BINS_NONE ? 0u : *in.n_records;
	if (ovf) return;  // the host grows the bins and re-issues the draw
	// candidates among the binned records: this tile's reference list, or (BINS_SCAN) every record once the tile's count is non-zero
	unsigned b0 = 0, b1 = 0;
	if (nrec) {
		if (in.bin_mode == BINS_LISTS) { b0 = in.bin_start[tile]; b1 = in.bin_start[tile + 1]; }
		else if (in.tile_cursor[tile]) b1 = nrec;
	}
	if (!touched && b0 == b1) return;
	const int tid = threadIdx.x;
	if (tid == 0) { s_clipped = 0u; s_nrec = 0u; s_ncand = 0u; }
	// 1. stage the tile's visibility keys in shared memory (and hand the global buffer back empty for the next draw)
	//    (all of a thread's loads first, then the stores: with the store inside the load loop the compiler keeps the loop rolled and a
	//    CTA starts with GT_PIX / TILE_THREADS dependent DRAM round trips instead of one)
	{
		constexpr int PER = GT_PIX / TILE_THREADS;
		unsigned long long kk[PER];
#pragma unroll
		for (int i = 0; i < PER; ++i) {
			const int p = tid + i * TILE_THREADS;
			const int px = x0 + (p & (GT - 1)), py = y0 + (p / GT);
			kk[i] = KEY_EMPTY;
			if (touched && px < fp.W && py >= fp.y_lo && py < fp.y_hi) kk[i] = in.vis[(size_t)py * fp.W + px];
		}
#pragma unroll
		for (int i = 0; i < PER; ++i) {
			const int p = tid + i * TILE_THREADS;
			const int px = x0 + (p & (GT - 1)), py = y0 + (p / GT);
			if (kk[i] != KEY_EMPTY) in.vis[(size_t)py * fp.W + px] = KEY_EMPTY;
			s_keys[p] = kk[i];
		}
	}
	__syncthreads();
	if (tid == 0) {
		in.tile_touched[tile] = 0u; in.tile_cursor[tile] = 0u;
		if (in.dirty) in.dirty[tile] = 1u;
	}
	// 2. binned triangles, BIN_ROUND candidates at a time: every thread loads and sets up ITS candidate (all gathers in flight at
	//    once) and the ones that reach into the tile are parked in shared memory; then warp w walks them over ITS strip of the tile
	//    (rows 4w .. 4w+3 with 8 warps), an 8x4 pixel block of the tile's own grid per step. A pixel belongs to one lane of one warp, so the running minimum
	//    needs no atomics, and a tile with a handful of large triangles (C1) keeps all warps busy instead of one.
	if (b1 > b0) {
		const int warp = tid >> 5, lane = tid & 31;
		const int lx = lane & 7, ly = lane >> 3;
		const int yb0 = max(y0, fp.y_lo), yb1 = min(min(y0 + GT, fp.H), fp.y_hi);
		constexpr int STRIP = GT / (TILE_THREADS / 32);  // rows per warp
		static_assert(STRIP >= 4 && STRIP % 4 == 0, "a warp's strip is made of 8x4 blocks");
		// BINS_SCAN: the candidates of a chunk of SCAN_CHUNK records are found first (box test only, a thread's record loads all in
		// flight), as indices in shared memory; BINS_LISTS: the tile's reference list is the one chunk.
		const bool scan = in.bin_mode == BINS_SCAN;
		for (unsigned cb = b0; cb < b1; cb += SCAN_CHUNK) {
			unsigned m0 = cb, m1 = b1;  // candidates [m0, m1) of this chunk: positions in in.items (lists) or in s_cand (scan)
			if (scan) {
#pragma unroll
				for (int k = 0; k < SCAN_CHUNK / TILE_THREADS; ++k) {
					const unsigned r = cb + k * TILE_THREADS + tid;
					if (r < b1) {
						const TriRecord t = in.records[r];
						// the pixel box of setup_triangle(), intersected with the tile
						const int X0 = max(x0, cvtt(floorf(min3f(t.x0, t.x1, t.x2)))), X1 = min(min(x0 + GT, fp.W), cvtt(ceilf(max3f(t.x0, t.x1, t.x2))));
						const int Y0 = max(yb0, cvtt(floorf(min3f(t.y0, t.y1, t.y2)))), Y1 = min(yb1, cvtt(ceilf(max3f(t.y0, t.y1, t.y2))));
						if (X0 < X1 && Y0 < Y1) s_cand[atomicAdd(&s_ncand, 1u)] = (unsigned short)(r - cb);
					}
				}
				__syncthreads();
				m0 = 0; m1 = s_ncand;
			}
			for (unsigned rb = m0; rb < m1; rb += BIN_ROUND) {
				const unsigned r = rb + tid;
				if (tid < BIN_ROUND && r < m1) {
					const TriRecord t = in.records[scan ? cb + s_cand[r] : in.items[r]];
					Setup s;
					if (setup_triangle(t.x0, t.y0, t.x1, t.y1, t.x2, t.y2, t.z0, t.z1, t.z2, fp.W, fp.y_lo, fp.y_hi, s)) {
						const int bx0 = max(s.X0, x0), bx1 = min(s.X1, x0 + GT), by0 = max(s.Y0, yb0), by1 = min(s.Y1, yb1);
						if (bx0 < bx1 && by0 < by1) {
							const unsigned at = atomicAdd(&s_nrec, 1u);
							float* d = s_rec[at];
							d[0] = s.a0; d[1] = s.b0; d[2] = s.c0; d[3] = s.a1; d[4] = s.b1; d[5] = s.c1; d[6] = s.a2; d[7] = s.b2; d[8] = s.c2;
							d[9] = s.inv_area; d[10] = s.z0; d[11] = s.z1; d[12] = s.z2;
							d[13] = __uint_as_float((unsigned)s.fminx); d[14] = __uint_as_float(t.ordinal);
							d[15] = __uint_as_float((unsigned)(bx0 - x0) | ((unsigned)(bx1 - x0) << 8) | ((unsigned)(by0 - y0) << 16) | ((unsigned)(by1 - y0) << 24));
						}
					}
				}
				__syncthreads();
				const unsigned m = s_nrec;
				const int sy0 = warp * STRIP, sy1 = sy0 + STRIP;  // this warp's rows, tile-relative
				for (unsigned i = 0; i < m; ++i) {
					const float* d = s_rec[i];
					const unsigned box = __float_as_uint(d[15]);
					const int qx0 = (int)(box & 255u), qx1 = (int)((box >> 8) & 255u);
					const int qy0 = max((int)((box >> 16) & 255u), sy0), qy1 = min((int)(box >> 24), sy1);
					if (qy0 >= qy1) continue;
					Setup q;
					q.a0 = d[0]; q.b0 = d[1]; q.c0 = d[2]; q.a1 = d[3]; q.b1 = d[4]; q.c1 = d[5]; q.a2 = d[6]; q.b2 = d[7]; q.c2 = d[8];
					q.inv_area = d[9]; q.z0 = d[10]; q.z1 = d[11]; q.z2 = d[12];
					q.fminx = (int)__float_as_uint(d[13]);
					const unsigned qord = __float_as_uint(d[14]);
					// blocks on the tile's own 8x4 grid: pixel (x, y) is always lane (x & 7, y & 3) of its strip's warp, whatever the record
					for (int ry = qy0 & ~3; ry < qy1; ry += 4)
						for (int rx = qx0 & ~7; rx < qx1; rx += 8) {
							const int tx_ = rx + lx, ty_ = ry + ly;
							if (tx_ < qx0 || tx_ >= qx1 || ty_ < qy0 || ty_ >= qy1) continue;
							const int px = x0 + tx_, py = y0 + ty_;
							float c0, c1, c2, al, be, ga;
							if (!coverage(q, px, py, c0, c1, c2)) continue;
							const float z = interp_z(q, c0, c1, c2, al, be, ga);
							if (!z_draws(z)) continue;
							const unsigned long long key = make_key(z, qord);
							if constexpr (PEEL) {
								if (!(key > in.floor[(size_t)py * fp.W + px])) continue;
							}
							unsigned long long* slot = &s_keys[ty_ * GT + tx_];
							if (key < *slot) *slot = key;
						}
				}
				__syncthreads();  // everybody is done with this round's candidates
				if (tid == 0) s_nrec = 0u;
				__syncthreads();
			}
			if (scan) {
				__syncthreads();  // (a chunk without candidates has no barrier between reading the count and this reset)
				if (tid == 0) s_ncand = 0u;
				__syncthreads();
			}
		}
	}
	// 3. deferred shading of the visible triangle of each pixel + framebuffer resolve.
	//    Not unrolled: one copy of the shading code keeps the kernel inside the instruction cache (unrolled x4 with prefetched
	//    indices: +12 % time; prefetched indices selected inside a rolled loop: +2 %).
	constexpr int PPT = GT_PIX / TILE_THREADS;
#if AXR_TILE_IDX_STASH
	// The vertex indices of all of a thread's pixels are fetched up front, together (one global round trip instead of one per pixel),
	// and parked in shared memory: slot p is written and read by the thread that owns pixel p, so no barrier is needed.
	__shared__ unsigned s_idx[3][GT_PIX];
#pragma unroll
	for (int i = 0; i < PPT; ++i) {
		const int blk = i * (TILE_THREADS / 32) + (tid >> 5);
#if AXR_QUAD_LANES
		const int lane = tid & 31, lx = ((lane >> 3) & 1) * 4 + (lane & 3), ly = (lane >> 4) * 2 + ((lane >> 2) & 1);
		const int p = in.row_major ? blk * GT + lane : ((blk >> 2) * 4 + ly) * GT + (blk & 3) * 8 + lx;
#else
		const int p = in.row_major ? blk * GT + (tid & 31) : ((blk >> 2) * 4 + ((tid & 31) >> 3)) * GT + (blk & 3) * 8 + (tid & 7);
#endif
		const unsigned long long k = s_keys[p];
		if (k == KEY_EMPTY) continue;
		const unsigned* ip = mesh.idx + (size_t)((unsigned)(k & 0xFFFFFFFFull) >> 3) * 3;
		s_idx[0][p] = __ldg(ip); s_idx[1][p] = __ldg(ip + 1); s_idx[2][p] = __ldg(ip + 2);
	}
#endif
#pragma unroll 1
	for (int i = 0; i < PPT; ++i) {
		// a warp = one compact 8x4 pixel block (neighbouring pixels share triangle vertices); the stores still fill whole 32 B
		// sectors (8 px x 4 B per row). Measured equal to 32x1 rows on C3.
		const int blk = i * (TILE_THREADS / 32) + (tid >> 5);
#if AXR_QUAD_LANES
		// quarter-warps (the unit a 16 B gather is served in) as 4x2 pixel blocks instead of 8x1 rows
		const int lane = tid & 31, lx = ((lane >> 3) & 1) * 4 + (lane & 3), ly = (lane >> 4) * 2 + ((lane >> 2) & 1);
		const int p = in.row_major ? blk * GT + lane : ((blk >> 2) * 4 + ly) * GT + (blk & 3) * 8 + lx;
#else
		const int p = in.row_major ? blk * GT + (tid & 31) : ((blk >> 2) * 4 + ((tid & 31) >> 3)) * GT + (blk & 3) * 8 + (tid & 7);
#endif
		const unsigned long long k = s_keys[p];
		const int px = x0 + (p & (GT - 1)), py = y0 + (p / GT);
		if (k == KEY_EMPTY) {
			if (FILL && px < fp.W && py >= fp.y_lo && py < fp.y_hi) {
				in.color[(size_t)py * fp.W + px] = in.fill_color;
				in.depth[(size_t)py * fp.W + px] = in.fill_depth;
			}
			continue;
		}
		const unsigned ord = (unsigned)(k & 0xFFFFFFFFull);
		// the framebuffer depth for the merge test does not depend on the triangle: issued first, so that a read that crosses PCIe
		// (host framebuffer, axr_draw_mesh_host) or NVLink has the index -> vertex gathers to hide behind
		const float fbz = FILL ? in.fill_depth : (in.read_depth ? in.depth_read[(size_t)py * fp.W + px] : INFINITY);
#if AXR_TILE_IDX_STASH
		const unsigned i0 = s_idx[0][p], i1 = s_idx[1][p], i2 = s_idx[2][p];
#elif AXR_IDX_PAD
		const uint4 iq = __ldg(mesh.idx4 + (ord >> 3));
		const unsigned i0 = iq.x, i1 = iq.y, i2 = iq.z;
#else
		const unsigned* ip = mesh.idx + (size_t)(ord >> 3) * 3;
		const unsigned i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2);
#endif
		const int res = shade_pixel<Shader, SMP, FAST>(mesh, u, fp, in, ord, i0, i1, i2, px, py, fbz);
		if (FILL && res != PIX_DONE) {  // not drawn (yet: k_shade_clipped may still store it)
			in.color[(size_t)py * fp.W + px] = in.fill_color;
			in.depth[(size_t)py * fp.W + px] = in.fill_depth;
		}
		if (res == PIX_CLIPPED) {
			// owned by a clipped face: the key goes back to the global buffer and the tile onto the list k_shade_clipped works through
			in.vis[(size_t)py * fp.W + px] = k;
			if (atomicExch(&s_clipped, 1u) == 0u) in.clip_tiles[atomicAdd(in.n_clip_tiles, 1u)] = (unsigned)tile;
			continue;
		}
		if constexpr (PEEL) {
			// discarded: the next pass looks for this pixel's next key above k. Otherwise the pixel is finished (the winner
			// was drawn, or it lost against the framebuffer and everything behind it would too): no key passes KEY_EMPTY.
			in.floor[(size_t)py * fp.W + px] = (res == PIX_DISCARDED) ? k : KEY_EMPTY;
			if (res == PIX_DISCARDED && *in.again == 0u) *in.again = 1u;
		}
	}
}

// Second, small kernel of the shading stage: the pixels k_tile_shade left behind because their visible triangle is a sub-triangle
// of a clipped face. A fixed grid walks the list of tiles that hold such pixels (empty for most frames: the CTAs then exit at once).
constexpr int CLIP_SHADE_THREADS = 128, CLIP_SHADE_CTAS = 148 * 2;
template <typename Shader, int SMP>
__global__ void __launch_bounds__(CLIP_SHADE_THREADS) k_shade_clipped(const __grid_constant__ MeshView mesh, const __grid_constant__ Uniforms u,
                                                                     const __grid_constant__ FrameParams fp, const __grid_constant__ TileIn in) {
	constexpr bool PEEL = Shader::DISCARDS;
	__shared__ unsigned long long s_key[GT_PIX];
	__shared__ unsigned short s_pix[GT_PIX];
	__shared__ unsigned s_n;
	const unsigned n = *in.n_clip_tiles;
	for (unsigned t = blockIdx.x; t < n; t += gridDim.x) {
		const int tile = (int)in.clip_tiles[t];
		const int x0 = (tile % fp.ntx) * GT, y0 = (tile / fp.ntx) * GT;
		if (threadIdx.x == 0) s_n = 0u;
		__syncthreads();
		// the keys k_tile_shade handed back (all loads of a thread in flight together), compacted so that every thread gets its share
#pragma unroll
		for (int i = 0; i < GT_PIX / CLIP_SHADE_THREADS; ++i) {
			const int p = i * CLIP_SHADE_THREADS + threadIdx.x;
			const int px = x0 + (p & (GT - 1)), py = y0 + (p / GT);
			if (px >= fp.W || py < fp.y_lo || py >= fp.y_hi) continue;
			unsigned long long* g = in.vis + (size_t)py * fp.W + px;
			const unsigned long long k = *g;
			if (k == KEY_EMPTY) continue;
			*g = KEY_EMPTY;
			const unsigned at = atomicAdd(&s_n, 1u);
			s_key[at] = k; s_pix[at] = (unsigned short)p;
		}
		__syncthreads();
		const unsigned m = s_n;
		for (unsigned i = threadIdx.x; i < m; i += CLIP_SHADE_THREADS) {
			const unsigned long long k = s_key[i];
			const int p = s_pix[i];
			const int px = x0 + (p & (GT - 1)), py = y0 + (p / GT);
			const bool discarded = shade_pixel_clipped<Shader, SMP>(mesh, u, fp, in, (unsigned)(k & 0xFFFFFFFFull), px, py);
			if constexpr (PEEL) {
				in.floor[(size_t)py * fp.W + px] = discarded ? k : KEY_EMPTY;
				if (discarded && *in.again == 0u) *in.again = 1u;
			} else {
				(void)discarded;
			}
		}
		__syncthreads();
	}
}

// ---- shader plug-ins (include/axr_shader_plugin.cuh): plug-in and library must have been built from the same kernel headers
// changes whenever a structure that crosses the plugin boundary changes size, or the launch shapes do
constexpr unsigned long long plugin_layout_hash() {
	unsigned long long h = 0xA11CE5ull;
	const unsigned long long parts[] = {sizeof(MeshView), sizeof(Uniforms), sizeof(FrameParams), sizeof(TileIn), sizeof(Material), (unsigned long long)TILE_THREADS,
	                                    (unsigned long long)GT, (unsigned long long)CLIP_SHADE_THREADS, (unsigned long long)BIN_ROUND, 4ull /* revision */};
	for (unsigned long long p : parts) h = (h ^ p) * 0x100000001B3ull;
	return h;
}

}  // namespace axr
