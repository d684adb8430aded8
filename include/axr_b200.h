/* axr_b200 — C ABI of the B200-native tiled rasterisation path (libaxr_b200.so).
 *
 * This is the drop-in boundary for AxiomR's hot path, AR::TiledPipeline::drawMesh
 * (reference src/tiled_pipeline.cpp:143-322) behind AR::Pipeline / AR::IShader
 * (reference include/pipeline.hpp:18-27, include/IShader.hpp:30-46). The reference has no FFI of
 * its own (static C++ linkage); each entry point below names the reference interface it replaces.
 * The C++ adapter that keeps the reference's class names on top of this ABI lives in
 * axiomr_b200/host/, the binding a maintainer adds is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types; no exceptions cross the boundary.
 *   - return 0 (AXR_OK) on success, negative axr_status on error; axr_last_error() has the text.
 *   - matrices are float[16] column-major exactly as glm::mat4 stores them (m[col*4+row]).
 *   - host buffers passed in are copied before the call returns (caller keeps ownership).
 *   - one context is used by one host thread at a time (the reference's drawMesh is not re-entrant either).
 *   - there is NO CPU fallback: every call needs a CUDA device (sm_100a); axr_create fails otherwise.
 *   - work is enqueued on the context's CUDA stream (plus an internal one for the geometry stages when axr_set_overlap is on);
 *     axr_sync / axr_resolve are the synchronisation points.
 */
#ifndef AXR_B200_H
#define AXR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AXR_ABI_VERSION 1

typedef enum axr_status {
	AXR_OK = 0,
	AXR_ERR_INVALID = -1,     /* bad argument / handle (reference: silent return, src/tiled_pipeline.cpp:146) */
	AXR_ERR_CUDA = -2,        /* CUDA runtime error, text in axr_last_error */
	AXR_ERR_NO_DEVICE = -3,   /* no sm_100 device: there is no CPU path */
	AXR_ERR_CAPACITY = -4,    /* mesh too large for 32-bit ordinals (reference: std::bad_alloc from the 16 MB arena, include/tiled_pipeline.hpp:98-104) */
	AXR_ERR_MATERIAL = -5,    /* shader needs a texture the material lacks (reference: null unique_ptr deref, include/shaders/shaders.hpp:178,210) */
	AXR_ERR_UNSUPPORTED = -6  /* unknown shader kind (reference: any IShader subclass; device functors exist for the shipped ones) */
} axr_status;

/* IShader implementations with a device functor (reference include/shaders/shaders.hpp:19-58, 136-250, 252-423).
 * AXR_SHADER_CUTOUT is not a reference shader: none of the shipped ones ever returns true (= discard) from fragment()
 * (include/IShader.hpp:38), so the discard branch of the raster loop (src/tiled_pipeline.cpp:571-577) is covered with an
 * alpha-tested Lambert shader written against the reference's IShader contract (oracle/ref_harness.cpp: CutoutShader;
 * needs the diffuse texture, discards where its alpha < 0.5). A draw with it is depth-peeled and synchronous. */
typedef enum axr_shader_kind { AXR_SHADER_FLAT = 0, AXR_SHADER_PHONG = 1, AXR_SHADER_PBR = 2, AXR_SHADER_CUTOUT = 3,
                               AXR_SHADER_PLUGIN_BASE = 64 /* + i: shaders added at run time, axr_load_shader_plugin */ } axr_shader_kind;

/* Texture::sample mode. NEAREST is the reference (include/texture.hpp:12-34). BILINEAR is an extension
 * (BASELINE.json config 3) with no reference counterpart; it is checked against oracle/axr_oracle.c only. */
typedef enum axr_sampler { AXR_SAMPLER_NEAREST = 0, AXR_SAMPLER_BILINEAR = 1 } axr_sampler;

/* Arithmetic of the shading stage. Coverage, depth and the choice of texels are bit-exact against the reference in both modes.
 * EXACT: colour arithmetic individually rounded in the reference's order (colour bytes identical to the CPU reference up to its
 * libm's powf). FAST (default): the same formulas with fused multiply-adds, the SFU reciprocal square root / log2 / exp2 and the
 * vertex stage applied once per pixel by linearity — 8-bit colour within 1 LSB of the reference (the tolerance BASELINE.json
 * states), about half the instructions per pixel. The environment variable AXR_B200_COLOR_MATH=exact|fast sets the default
 * a context is created with. Shaders that discard (AXR_SHADER_CUTOUT) always run EXACT: their colour test decides coverage. */
typedef enum axr_color_math { AXR_COLOR_EXACT = 0, AXR_COLOR_FAST = 1 } axr_color_math;

/* Public shader parameters (FlatShader::lightDirection; PhongShader/PBRShader::lightDirection, lightColor). */
typedef struct axr_shader_params {
	float light_dir[3];
	float light_color[3];
} axr_shader_params;

typedef struct axr_config {
	int device;            /* CUDA device ordinal */
	int width, height;     /* Framebuffer(width, height, useDepth=true), reference include/framebuffer.hpp:12 */
	int sampler;           /* axr_sampler */
	/* Multi-GPU screen-space band owned by this context: pixel rows [band_y0, band_y1). 0,0 = whole frame.
	 * Rows outside the band are never rasterised, shaded or written. band_y0 must be a multiple of 16
	 * (the reference tile size, include/tiled_pipeline.hpp:28) so per-pixel results do not depend on the split. */
	int band_y0, band_y1;
	void* stream;          /* optional cudaStream_t to enqueue on (e.g. the caller's timing stream); NULL = own stream */
	uint64_t reserved[4];  /* must be zero */
} axr_config;

typedef struct axr_ctx axr_ctx;
typedef int32_t axr_mesh;  /* handles are small non-negative integers */
typedef int32_t axr_tex;
#define AXR_NO_TEXTURE (-1)

/* MaterialGroup {materialName, startIndex, faceCount} (reference include/mesh.hpp:35-39) without the name. */
typedef struct axr_group {
	uint64_t first_face;
	uint64_t face_count;
} axr_group;

/* Counters of the most recent draw (diagnostics / tests / bench launch accounting). */
typedef struct axr_stats {
	uint64_t faces;            /* input faces */
	uint64_t clipped_faces;    /* faces that took the clip slow path (some vertex outside some plane) */
	uint64_t triangles;        /* post-clip, front-facing triangles (|m_Triangles| in the reference) */
	uint64_t small_triangles;  /* rasterised directly by the setup kernel */
	uint64_t binned_triangles; /* setup records sent through tile bins */
	uint64_t bin_refs;         /* (triangle, tile) references written */
	uint64_t kernel_launches;  /* CUDA kernels launched by this draw */
	uint64_t redo;             /* 1 if the draw was re-issued after growing bin capacity */
} axr_stats;

/* ---- lifetime: replaces `new Framebuffer(w,h,true)` + `new TiledPipeline(threads, camera, fb)`
 *      (reference src/renderer.cpp:66-71). The thread-count argument has no meaning on the GPU. */
int axr_create(const axr_config* cfg, axr_ctx** out);
void axr_destroy(axr_ctx* ctx);
const char* axr_last_error(const axr_ctx* ctx);  /* ctx may be NULL: error of the last failed axr_create */
int axr_abi_version(void);

/* ---- scene data: replaces Mesh::getVertices()/getFaces()/getMaterialGroups()/getMaterial() consumption
 *      (reference src/tiled_pipeline.cpp:157-159,176-179) and Texture(path) data (reference src/texture.cpp:21-36).
 *      vertices: n_verts x 14 f32 in AR::Vertex layout (reference include/mesh.hpp:9-18), 56-byte stride.
 *      indices: 3 u32 per face (reference asserts triangles, src/tiled_pipeline.cpp:202).
 *      groups may be NULL: one group covering every face. Groups are contiguous ascending face ranges ending at n_faces; the first
 *      may start after face 0: the faces in front of it belong to no group and are not drawn (an OBJ's faces before its first `usemtl`). */
int axr_upload_mesh(axr_ctx* ctx, const float* vertices, uint64_t n_verts, const uint32_t* indices,
                    uint64_t n_faces, const axr_group* groups, uint32_t n_groups, axr_mesh* out);
int axr_free_mesh(axr_ctx* ctx, axr_mesh mesh);
/* The reference reads its host Mesh on every drawMesh; here a mesh is uploaded once. A caller that edits vertices in place (same
 * vertex count, same faces: animation, morphing) re-sends them with this call instead of freeing and re-uploading. */
int axr_update_mesh_vertices(axr_ctx* ctx, axr_mesh mesh, const float* vertices, uint64_t n_verts);
/* rgba: w*h*4 bytes, row 0 = image top — what stbi_load(..., STBI_rgb_alpha) returns. */
int axr_upload_texture(axr_ctx* ctx, const uint8_t* rgba, int w, int h, axr_tex* out);
int axr_free_texture(axr_ctx* ctx, axr_tex tex);
/* Material of one group: diffuse/bump/metallic/roughness/ao textures + specularExponent ("Ns")
 * (reference include/mesh.hpp:20-34). Pass AXR_NO_TEXTURE for absent maps. */
int axr_set_material(axr_ctx* ctx, axr_mesh mesh, uint32_t group, axr_tex diffuse, axr_tex bump, axr_tex metallic,
                     axr_tex roughness, axr_tex ao, float specular_exponent);

/* ---- mesh ingestion: replaces `AR::Mesh(path)` (reference src/mesh.cpp:8-27): parseModelFile (:300-415 — OBJ text, value
 *      de-duplication of (position, uv, normal) in first-occurrence order, fan triangulation, one material group per `usemtl`),
 *      then calculateTangentBitangent (:222-298, on the device) and the upload. The arrays equal the reference loader's bit for bit
 *      (axr_mesh_read returns them: n_verts x 14 f32 in AR::Vertex layout, 3 u32 per face). As in the reference, faces in front of
 *      the first `usemtl` belong to no group and are not drawn (drawMesh walks the groups, src/tiled_pipeline.cpp:176-179); an OBJ
 *      without `usemtl` draws nothing. A face index without digits makes the reference's std::stoi throw; here it is AXR_ERR_INVALID.
 *      Materials stay with the caller (textures are decoded by stb_image in the reference): axr_parse_mtl lists, per `newmtl`, the
 *      name, Ns and the texture paths of the five maps the shaders read (loadMaterial / parseMaterialData, :65-220; a later entry
 *      with the same name replaces an earlier one, as in the reference's map); feed them to axr_upload_texture / axr_set_material. */
typedef struct axr_obj_info {
	uint64_t n_verts, n_faces;
	uint32_t n_groups, reserved;
	uint64_t first_drawn_face;   /* == n_faces when the OBJ has no `usemtl` */
} axr_obj_info;
typedef struct axr_mtl_entry {
	char name[128];
	float specular_exponent;     /* Ns */
	int has_map[5];              /* diffuse (map_Kd), bump (map_Bump | bump | norm), metallic (map_Ks | refl), roughness (map_Ns), ao (map_A0) */
	char map[5][512];            /* path as written in the file, blanks trimmed; relative to the MTL's directory */
} axr_mtl_entry;
int axr_load_obj(axr_ctx* ctx, const char* obj_text, size_t len, axr_mesh* out, axr_obj_info* info /* may be NULL */);
int axr_load_obj_file(axr_ctx* ctx, const char* path, axr_mesh* out, axr_obj_info* info);
int axr_mesh_group_info(axr_ctx* ctx, axr_mesh mesh, uint32_t group, char* name, size_t name_cap, uint64_t* first_face, uint64_t* face_count);
int axr_mesh_read(axr_ctx* ctx, axr_mesh mesh, float* vertices /* n_verts x 14, may be NULL */, uint32_t* indices /* may be NULL */);
int axr_parse_mtl(const char* mtl_text, size_t len, axr_mtl_entry* out, uint32_t cap, uint32_t* n_out /* entries in the file */);

/* ---- mesh ingestion helper: Mesh::calculateTangentBitangent (reference src/mesh.cpp:222-298) on the device.
 *      in: n_verts x 8 f32 (position3, uv2, normal3) as the OBJ parser leaves them + 3 u32 per face;
 *      out: n_verts x 14 f32 in AR::Vertex layout with the reference's tangents / bitangents, bit for bit (per-vertex sums are
 *      taken in face order like the reference's). Host in, host out, synchronous. */
int axr_generate_tangents(axr_ctx* ctx, const float* pos_uv_normal, uint64_t n_verts, const uint32_t* indices, uint64_t n_faces,
                          float* vertices_out);

/* ---- per-frame state: replaces Pipeline::setCamera / the uniform writes at src/tiled_pipeline.cpp:148-155.
 *      view_proj = Camera::getViewProjectionMatrix(), viewport = Camera::getViewportMatrix() (carried, unused by
 *      the shipped shaders), cam_pos = Camera::getPosition(). mvp = view_proj * model is formed per draw in glm order. */
int axr_set_uniforms(axr_ctx* ctx, const float view_proj[16], const float viewport[16], const float cam_pos[3]);
/* replaces Pipeline::setShader(IShader*) (reference src/pipeline.cpp:22-24); params = the shader's public fields. */
int axr_set_shader(axr_ctx* ctx, int kind, const axr_shader_params* params, size_t params_size);
/* A further IShader subclass at run time (reference include/IShader.hpp:30-46 — Pipeline::setShader takes any). On the device a shader is
 * a functor the tile kernels are instantiated with; its author writes it against include/axr_shader_plugin.cuh (the same two
 * entry points as the plugin contract: vertex() and fragment(), true = discard) and compiles it with nvcc into a shared library
 * (tools/build_shader_plugin.py). axr_load_shader_plugin opens it, checks that it was built against this library's kernel headers and
 * returns the shader kind to pass to axr_set_shader (AXR_SHADER_PLUGIN_BASE + i). A plug-in that declares DISCARDS is depth-peeled like
 * AXR_SHADER_CUTOUT; plug-ins always run the individually rounded (EXACT) colour arithmetic. axr_set_shader_user passes up to 8 floats
 * to the functor (Uniforms::user) beside the light direction / colour of axr_shader_params. */
int axr_load_shader_plugin(axr_ctx* ctx, const char* path, int* kind_out);
int axr_set_shader_user(axr_ctx* ctx, const float* values, uint32_t n);
int axr_set_sampler(axr_ctx* ctx, int sampler);
int axr_set_color_math(axr_ctx* ctx, int mode);  /* axr_color_math */

/* ---- framebuffer: replaces Framebuffer::clearColor / clearDepth (reference src/framebuffer.cpp:26-42) and the raw
 *      getColorData()/getDepthData() accessors (reference include/framebuffer.hpp:45-48). Colour bytes are B,G,R,A;
 *      row 0 is the bottom of the image (y up), depth is f32 NDC z with +inf = empty. */
int axr_clear(axr_ctx* ctx, uint32_t packed_argb, float depth);
int axr_upload_framebuffer(axr_ctx* ctx, const uint8_t* bgra, const float* depth);
/* Same, but only enqueued: the host buffers must stay untouched until the next axr_resolve / axr_sync (pinned memory, see
 * axr_host_alloc, is needed for the copy to be truly asynchronous). With axr_set_overlap(1) the geometry stages of the
 * following axr_draw_mesh run while the copy is still in flight (they do not read the framebuffer). */
int axr_upload_framebuffer_async(axr_ctx* ctx, const uint8_t* bgra, const float* depth);
/* Device -> host copy of the whole frame (or this context's band rows only, at their place in the full image),
 * synchronous: on return the draw(s) are complete, like the reference's drawMesh. NULL pointers are skipped. */
int axr_resolve(axr_ctx* ctx, uint8_t* bgra_out, float* depth_out);

/* ---- the hot path: replaces TiledPipeline::drawMesh(const glm::mat4& model, const Mesh&)
 *      (reference src/tiled_pipeline.cpp:143-322). Composites onto the current framebuffer contents with the
 *      reference's strict depth test; never clears. Asynchronous on the context stream. */
int axr_draw_mesh(axr_ctx* ctx, axr_mesh mesh, const float model[16]);
/* drawMesh with the reference's exact calling convention: composite onto a HOST framebuffer (Framebuffer::getColorData() /
 * getDepthData()), complete on return. Nothing is uploaded: the host arrays are mapped into the device's address space (pinned
 * memory from axr_host_alloc is mapped already, other memory is page-locked with cudaHostRegister on first use and remembered); the
 * merge test reads the host depth of the visible pixels through the mapping (128 B row reads over PCIe) and the pixels that pass
 * are stored by the tile kernel straight into the host arrays, 8 bytes per updated pixel. Falls back to upload / draw / resolve
 * when the host memory cannot be mapped. The device-resident framebuffer of the context is left unspecified by this call.
 * AXR_B200_HOST_DEPTH_ZEROCOPY=0 (read once at axr_create) uploads the whole host depth plane in row chunks instead of reading it
 * through the mapping (measured slower on C3: 0.98 against 0.87 ms per call). */
int axr_draw_mesh_host(axr_ctx* ctx, axr_mesh mesh, const float model[16], uint8_t* bgra, float* depth);
/* A buffer that axr_draw_mesh_host page-locked stays locked until axr_destroy; a caller that frees such a buffer earlier (a
 * std::vector, a numpy array) releases it first. A remembered registration is only reused when it contains the whole framebuffer;
 * one that merely overlaps it (freed memory whose address was reused) is dropped and taken again. Synchronises the context. */
int axr_host_release(axr_ctx* ctx, void* host_ptr);
int axr_sync(axr_ctx* ctx);
int axr_get_stats(axr_ctx* ctx, axr_stats* out);

/* ---- per-kernel device timing (CUDA events on the context stream, recorded around each kernel of every draw while
 *      enabled). Stage order: 0 vertex_xform, 1 setup_raster, 2 scan_tiles, 3 bin_scatter, 4 tile_shade.
 *      axr_get_kernel_times synchronises, returns the accumulated milliseconds per stage and the number of draws
 *      accumulated, and resets the accumulators. With axr_set_overlap(1) consecutive draws overlap, so the per-stage times
 *      then add up to more than the wall time of a sequence. */
#define AXR_NUM_STAGES 5
int axr_set_profiling(axr_ctx* ctx, int enabled);
int axr_get_kernel_times(axr_ctx* ctx, float ms_out[AXR_NUM_STAGES], uint64_t* draws_out);

/* ---- FP32 issue micro-benchmark: measured warp-instructions per second of this GPU for (0) separate
 *      FMUL + FADD, which is what the path executes (no contraction, for parity), and (1) FFMA. One warp-instruction = 32 lanes. */
int axr_measure_fp32_issue(axr_ctx* ctx, double* fmul_fadd_winst_per_s, double* ffma_winst_per_s);
/* The same for the other bound of the shading stage: data-dependent 16-byte gathers, one 32-byte DRAM sector each (vertex records,
 * attributes, texels of neighbouring pixels share little at one triangle per pixel). Measures how many randomly placed sectors per
 * second this GPU delivers (1 GiB buffer, 8 loads in flight per thread, 2368 CTAs): x 32 bytes = the random-sector bandwidth the
 * stage's DRAM traffic is held against (bench.py: roofline_gather). */
int axr_measure_gather(axr_ctx* ctx, double* sectors_per_s);

/* ---- pinned host memory for framebuffers / staging (cudaHostAlloc): makes axr_upload_framebuffer / axr_resolve
 *      run at full PCIe rate. Plain malloc'ed memory works too, just slower. */
void* axr_host_alloc(size_t bytes);
void axr_host_free(void* p);

/* ---- interop for callers that keep data on the device (bench timing, NCCL / peer composite) */
void* axr_stream(axr_ctx* ctx);                                          /* cudaStream_t */
int axr_framebuffer_device(axr_ctx* ctx, void** bgra_dev, void** depth_dev);  /* full-frame device pointers (W*H*4, W*H*4 bytes) */
/* Redirect this context's band output into another allocation laid out as a full frame (e.g. GPU 0's framebuffer
 * mapped through CUDA IPC / peer access): the resolve stores then go straight over NVLink. NULL restores the own buffers. */
int axr_set_output(axr_ctx* ctx, void* bgra_dev, void* depth_dev);
/* Dirty-tile tracking for composite targets that are re-cleared every frame (multi-GPU views / bands: GPU 0 owns the targets, the
 * other GPUs store into them over NVLink). With a dirty map set, every draw flags the 32x32-pixel tiles it may store into
 * (axr_dirty_map_entries() 32-bit flags, row-major tiles, ceil(W/32) per row; the map may live on another GPU);
 * axr_clear_dirty_tiles then clears only the flagged tiles of `count` targets laid out back to back (colour planes of W*H words, depth
 * planes of W*H floats, maps of axr_dirty_map_entries() flags, each group contiguous) and resets the flags — the clear of a sparsely
 * covered frame touches a fraction of it. stream: cudaStream_t to launch on, NULL = the context stream. */
int axr_dirty_map_entries(const axr_ctx* ctx);
int axr_set_dirty_map(axr_ctx* ctx, void* dirty_dev);
int axr_clear_dirty_tiles(axr_ctx* ctx, void* bgra_dev, void* depth_dev, void* dirty_dev, int count, uint32_t packed_argb, float depth, void* stream);
/* The same targets without clearing what the next frame overwrites anyway. With axr_set_output_fill on, a draw overwrites EVERY pixel of
 * the 32x32 tiles it touches: the shaded colour where a triangle is visible, (packed_argb, depth) elsewhere, and its merge test sees
 * `depth` instead of reading the target — valid for a target that holds exactly one draw on top of a clear to those values, which is
 * what a composite slot is (not for shaders that discard, nor for axr_draw_mesh_host). The owner of the targets then keeps two dirty maps
 * per target, alternating per use: axr_clear_stale_tiles clears only the tiles flagged in `prev` (the previous use) and not in `now`
 * (this use) — none at all while the camera stands still — and hands `prev` back all zero. Same layouts as axr_clear_dirty_tiles.
 * Two small kernels (one thread per (target, tile) lists the stale tiles, a fixed grid clears the listed ones) that share a list owned by
 * the context: calls on one context have to be ordered (one stream, or events between them). */
int axr_set_output_fill(axr_ctx* ctx, int enabled, uint32_t packed_argb, float depth);
/* Pixel -> lane mapping of the shading stage for outputs behind a link (axr_set_output into another GPU's memory): with rows on, a warp
 * shades one 32 x 1 pixel row instead of an 8 x 4 block, so its colour and depth stores are 128 contiguous bytes each — NVLink moves
 * those at about twice the rate of the 32-byte pieces a block row makes (what axr_draw_mesh_host does for PCIe). Results are identical. */
int axr_set_output_rows(axr_ctx* ctx, int enabled);
int axr_clear_stale_tiles(axr_ctx* ctx, void* bgra_dev, void* depth_dev, void* dirty_prev_dev, void* dirty_now_dev, int count, uint32_t packed_argb,
                          float depth, void* stream);
/* Overlap consecutive draws: with overlap on, the geometry stages (vertex, setup, bins) of draw i+1 are enqueued on a second,
 * higher-priority stream and run beside the tile / shading kernel of draw i (two sets of per-draw buffers alternate). Results
 * are identical; throughput of back-to-back draws rises by a few percent (C3: 0.464 -> 0.433 ms per frame). Default: off, which
 * keeps every kernel of a draw on the context stream in launch order (what per-kernel timings and profiles assume). */
int axr_set_overlap(axr_ctx* ctx, int enabled);
/* The tile kernel reads the output depth for the reference's merge test `z < fbZ` (src/tiled_pipeline.cpp:1148-1156). When the
 * caller guarantees that the output was just cleared to depth = +inf and receives exactly one draw (e.g. a per-view slot on
 * another GPU, where that read would cross NVLink), the read can be turned off: every drawable z passes `z < +inf`.
 * Default: enabled. */
int axr_set_depth_read(axr_ctx* ctx, int enabled);
/* Device memory that other processes can map (cudaMalloc + cudaIpcGetMemHandle), zero-filled: composite targets + dirty maps on GPU 0. */
int axr_alloc_shared(axr_ctx* ctx, size_t bytes, void** dev_ptr_out, void* handle64_out);
int axr_free_shared(axr_ctx* ctx, void* dev_ptr);
/* CUDA IPC handles (64 bytes each) of the own framebuffer allocations, for one-process-per-GPU compositing. */
int axr_framebuffer_ipc(axr_ctx* ctx, void* color_handle64, void* depth_handle64);
int axr_open_ipc(axr_ctx* ctx, const void* handle64, void** dev_ptr_out);
int axr_close_ipc(axr_ctx* ctx, void* dev_ptr);

#ifdef __cplusplus
}
#endif
#endif /* AXR_B200_H */
