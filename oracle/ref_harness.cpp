// TEST INFRASTRUCTURE — not part of the product path.
//
// Headless C-ABI harness around the UNMODIFIED AxiomR reference sources (compiled where they lie
// under /root/reference by oracle/Makefile, output into oracle/_ref/libaxr_ref.so). It drives the
// reference's own AR::TiledPipeline::drawMesh (reference src/tiled_pipeline.cpp:143-322), IShader
// implementations (include/shaders/shaders.hpp), Texture::sample (include/texture.hpp:12-34),
// Pipeline::clipTriangle (src/pipeline.cpp:176-228) and Camera (src/camera.cpp) on caller-supplied
// arrays, so tests can (1) pin the C restatement in oracle/axr_oracle.c bit-for-bit and
// (2) generate the golden fixtures under tests/golden/. bench.py times it as cpu_baseline kind
// "reference". Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
//
// Access to private members (Mesh arrays, Camera matrices) is obtained by re-declaring the access
// specifiers for THIS translation unit only; the reference TUs themselves are compiled unmodified,
// and GCC's layout does not depend on access specifiers.
#include "prelude.hpp"
#include <chrono>
#include <cstdio>
#include <map>
#include <mutex>
#include <unistd.h>

#define private public
#define protected public
#include "tiled_pipeline.hpp"
#include "shaders/shaders.hpp"
#include "camera.hpp"
#include "framebuffer.hpp"
#include "mesh.hpp"
#include "texture.hpp"
#undef private
#undef protected

// Link stubs: src/camera.cpp:102-126 references these Win32-backed methods (never called here).
namespace AR {
bool Window::isMouseButtonDown(int) const { return false; }
bool Window::isKeyDown(char) const { return false; }
}

namespace {

struct TexFile {
	std::string path;
	~TexFile() { if (!path.empty()) unlink(path.c_str()); }
};

// Writes RGBA8 (row 0 = image top, the layout stbi_load returns, reference src/texture.cpp:24)
// as an uncompressed top-left-origin 32-bit TGA so it enters through the reference's own
// Texture(path) constructor and stb_image decoder.
bool write_tga(const std::string& path, const uint8_t* rgba, int w, int h) {
	FILE* f = fopen(path.c_str(), "wb");
	if (!f) return false;
	uint8_t hdr[18] = {0};
	hdr[2] = 2;
	hdr[12] = (uint8_t)(w & 255); hdr[13] = (uint8_t)(w >> 8);
	hdr[14] = (uint8_t)(h & 255); hdr[15] = (uint8_t)(h >> 8);
	hdr[16] = 32; hdr[17] = 0x28;  // 8 alpha bits, top-left origin
	fwrite(hdr, 1, 18, f);
	std::vector<uint8_t> row((size_t)w * 4);
	for (int y = 0; y < h; ++y) {
		const uint8_t* s = rgba + (size_t)y * w * 4;
		for (int x = 0; x < w; ++x) {
			row[x * 4 + 0] = s[x * 4 + 2]; row[x * 4 + 1] = s[x * 4 + 1];
			row[x * 4 + 2] = s[x * 4 + 0]; row[x * 4 + 3] = s[x * 4 + 3];
		}
		fwrite(row.data(), 1, row.size(), f);
	}
	fclose(f);
	return true;
}

std::unique_ptr<AR::Texture> make_texture(const uint8_t* rgba, int w, int h, int slot) {
	if (!rgba || w <= 0 || h <= 0) return nullptr;
	TexFile tf;
	char name[256];
	snprintf(name, sizeof name, "/tmp/axr_ref_tex_%d_%d.tga", (int)getpid(), slot);
	tf.path = name;
	if (!write_tga(tf.path, rgba, w, h)) return nullptr;
	auto t = std::make_unique<AR::Texture>(tf.path);
	if (!t->m_Data || t->m_Width != w || t->m_Height != h) return nullptr;
	return t;
}

// A fourth IShader, written for this harness: the reference ships no shader that returns true (= discard) from fragment(),
// so its discard branch (reference src/tiled_pipeline.cpp:571-577) would otherwise have no parity target. Alpha-tested
// Lambert: FlatShader's normal handling (include/shaders/shaders.hpp:26-57) plus the diffuse texel; fragments whose texel
// alpha is below 0.5 are discarded. It runs inside the UNMODIFIED reference pipeline through the IShader plugin contract
// (include/IShader.hpp:30-46); oracle/axr_oracle.c (kind 3) and the CUDA CutoutShader functor restate it.
struct CutoutShader : public AR::IShader {
	glm::vec3 lightDirection;
	AR::VertexOutput vertex(const AR::Vertex& v, int) override {
		AR::VertexOutput o;
		o.uv = v.uv;
		o.normal = glm::mat3(glm::transpose(glm::inverse(model))) * v.normal;
		return o;
	}
	bool fragment(glm::vec3& bar, glm::vec4& color, const AR::VSTransformedTriangle& tri) override {
		glm::vec2 uv = bar.x * tri[0].uv + bar.y * tri[1].uv + bar.z * tri[2].uv;
		glm::vec4 texel = material->diffuseTexture->sample(uv);
		if (texel.a < 0.5f) return true;
		glm::vec3 n = bar.x * tri[0].normal + bar.y * tri[1].normal + bar.z * tri[2].normal;
		n = glm::normalize(n);
		float intensity = std::clamp(glm::dot(-lightDirection, n), 0.0f, 1.0f);
		color = glm::vec4(glm::vec3(texel) * intensity, 1.0f);
		return false;
	}
};

glm::mat4 to_mat4(const float* m) {
	glm::mat4 r;
	for (int c = 0; c < 4; ++c) r[c] = glm::vec4(m[c * 4 + 0], m[c * 4 + 1], m[c * 4 + 2], m[c * 4 + 3]);
	return r;
}
void from_mat4(const glm::mat4& m, float* out) {
	for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) out[c * 4 + r] = m[c][r];
}

}  // namespace

extern "C" {

struct axr_ref_scene {
	int width, height;
	int threads;             // TiledPipeline(threads, ...) (reference src/renderer.cpp:71 passes hardware_concurrency)
	int shader_kind;         // 0 FlatShader, 1 PhongShader, 2 PBRShader, 3 CutoutShader (above; needs the diffuse texture)
	float light_dir[3];
	float light_color[3];
	float specular_exponent; // Material::specularExponent ("Ns")
	const uint8_t* tex[5];   // diffuse, bump, metallic, roughness, ao : RGBA8, row 0 = top; may be null
	int tex_w[5], tex_h[5];
	float view_proj[16];     // column-major, injected into Camera::m_ViewProjection
	float cam_pos[3];
	float model[16];
	int chunk_faces;         // faces per drawMesh call (<= 30000: 16 MB triangle arena, include/tiled_pipeline.hpp:98-104)
};

int axr_ref_hardware_concurrency(void) { return (int)std::thread::hardware_concurrency(); }

// Camera matrices from the reference's own Camera (ctor + setViewport + update(0), src/camera.cpp).
int axr_ref_camera(const float pos[3], const float target[3], float fov_deg, float aspect, int w, int h,
                   float view_proj_out[16], float viewport_out[16]) {
	AR::Camera cam(glm::vec3(pos[0], pos[1], pos[2]), glm::vec3(target[0], target[1], target[2]), fov_deg, aspect);
	cam.setViewport(0, 0, w, h);
	cam.update(0.0f);
	from_mat4(cam.getViewProjectionMatrix(), view_proj_out);
	from_mat4(cam.getViewportMatrix(), viewport_out);
	return 0;
}

// mvp = viewProj * model exactly as reference src/tiled_pipeline.cpp:149 computes it.
void axr_ref_mat4_mul(const float a[16], const float b[16], float out[16]) { from_mat4(to_mat4(a) * to_mat4(b), out); }

// One AR::TiledPipeline per thread count, kept for the life of the process — like the application, which builds its pipeline once
// (reference src/renderer.cpp:71) and never destroys it before exit. Destroying one right after a draw can hang: AR::ThreadPool's
// destructor (reference src/thread_pool.cpp:24-29) sets `stop` and notifies WITHOUT holding the queue mutex, so a worker that has
// just evaluated the wait predicate and not yet blocked misses the wake-up and the join never returns (observed: three of three
// fuzz campaigns of ~10^4 short renders each ended in that futex wait). The reference sources are not ours to change, so the
// harness simply does not run that destructor on the fast path; when a pipeline has to be replaced (arena overflow below) it first
// gives the idle workers time to park.
static std::mutex g_pipes_mutex;
static std::map<int, AR::TiledPipeline*> g_pipes;

static AR::TiledPipeline* pipeline_for(int threads, AR::Camera* cam, AR::Framebuffer* fb, AR::IShader* shader, bool replace) {
	std::lock_guard<std::mutex> lock(g_pipes_mutex);
	AR::TiledPipeline*& p = g_pipes[threads];
	if (p && replace) {
		std::this_thread::sleep_for(std::chrono::milliseconds(5));
		delete p;
		p = nullptr;
	}
	if (!p) p = new AR::TiledPipeline((size_t)threads, cam, fb);  // heap: the object embeds a 32 MB arena (include/tiled_pipeline.hpp:103)
	p->setCamera(cam);
	p->setFramebuffer(fb);
	p->setShader(shader);
	return p;
}

// Renders `n_faces` triangles through the reference's TiledPipeline in chunks, compositing onto
// color_inout (BGRA8) / depth_inout (f32) exactly as consecutive Renderer::drawMesh calls would.
// seconds_out (optional) receives the wall time spent inside drawMesh calls only.
int axr_ref_render(const axr_ref_scene* sc, const float* vertices, uint64_t n_verts, const uint32_t* indices,
                   uint64_t n_faces, uint8_t* color_inout, float* depth_inout, double* seconds_out) {
	try {
		const int W = sc->width, H = sc->height;
		AR::Framebuffer fb(W, H, true);
		std::memcpy(fb.getColorData(), color_inout, (size_t)W * H * 4);
		std::memcpy(fb.getDepthData(), depth_inout, (size_t)W * H * sizeof(float));

		AR::Camera cam(glm::vec3(0, 0, 5), glm::vec3(0, 0, 0), 60.0f, (float)W / (float)H);
		cam.setViewport(0, 0, W, H);
		cam.update(0.0f);
		cam.m_ViewProjection = to_mat4(sc->view_proj);
		cam.m_Position = glm::vec3(sc->cam_pos[0], sc->cam_pos[1], sc->cam_pos[2]);

		AR::FlatShader flat;
		AR::PhongShader phong;
		AR::PBRShader pbr;
		glm::vec3 L(sc->light_dir[0], sc->light_dir[1], sc->light_dir[2]);
		glm::vec3 LC(sc->light_color[0], sc->light_color[1], sc->light_color[2]);
		flat.lightDirection = L;
		phong.lightDirection = L; phong.lightColor = LC;
		pbr.lightDirection = L; pbr.lightColor = LC;
		CutoutShader cutout;
		cutout.lightDirection = L;
		AR::IShader* shader = sc->shader_kind == 0 ? (AR::IShader*)&flat
		                    : sc->shader_kind == 1 ? (AR::IShader*)&phong
		                    : sc->shader_kind == 2 ? (AR::IShader*)&pbr : (AR::IShader*)&cutout;

		const int threads = std::max(1, sc->threads);
		AR::TiledPipeline* pipe = pipeline_for(threads, &cam, &fb, shader, false);

		AR::Mesh mesh;
		mesh.m_Vertices.resize(n_verts);
		static_assert(sizeof(AR::Vertex) == 56, "Vertex layout");
		std::memcpy((void*)mesh.m_Vertices.data(), vertices, n_verts * sizeof(AR::Vertex));
		auto mat = std::make_unique<AR::Material>();
		mat->name = "m0";
		mat->specularExponent = sc->specular_exponent;
		mat->diffuseTexture = make_texture(sc->tex[0], sc->tex_w[0], sc->tex_h[0], 0);
		mat->bumpTexture = make_texture(sc->tex[1], sc->tex_w[1], sc->tex_h[1], 1);
		mat->metallicTexture = make_texture(sc->tex[2], sc->tex_w[2], sc->tex_h[2], 2);
		mat->roughnessTexture = make_texture(sc->tex[3], sc->tex_w[3], sc->tex_h[3], 3);
		mat->aoTexture = make_texture(sc->tex[4], sc->tex_w[4], sc->tex_h[4], 4);
		if ((sc->shader_kind == 1 || sc->shader_kind == 2) && (!mat->diffuseTexture || !mat->bumpTexture)) return -2;
		if (sc->shader_kind == 3 && !mat->diffuseTexture) return -2;
		if (sc->shader_kind == 2 && (!mat->metallicTexture || !mat->roughnessTexture || !mat->aoTexture)) return -2;
		mesh.m_Materials["m0"] = std::move(mat);

		const glm::mat4 model = to_mat4(sc->model);
		uint64_t chunk = (uint64_t)std::max(1, sc->chunk_faces);
		double secs = 0.0;
		for (uint64_t f0 = 0; f0 < n_faces;) {
			const uint64_t n = std::min(chunk, n_faces - f0);
			mesh.m_Faces.resize(n);
			for (uint64_t i = 0; i < n; ++i) {
				const uint32_t* ix = indices + (f0 + i) * 3;
				mesh.m_Faces[i].vertexIndices.assign(ix, ix + 3);
			}
			mesh.m_MaterialGroups.clear();
			mesh.m_MaterialGroups.push_back({"m0", 0, (size_t)n});
			auto t0 = std::chrono::steady_clock::now();
			try {
				pipe->drawMesh(model, mesh);
			} catch (const std::bad_alloc&) {
				// The 16 MB triangle arena (include/tiled_pipeline.hpp:98-104) overflowed: how many post-cull triangles fit depends
				// on the thread count and on how many faces of the chunk survive culling. The throw happens while m_Triangles is
				// being assembled, before anything is rasterised or merged, so the chunk can be redrawn: fresh pipeline (the
				// monotonic arena never gives memory back), half the chunk. Time of the failed attempt is not counted.
				if (chunk == 1) throw;
				chunk = std::max<uint64_t>(1, chunk / 2);
				pipe = pipeline_for(threads, &cam, &fb, shader, true);
				continue;
			}
			secs += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			f0 += n;
		}
		std::memcpy(color_inout, fb.getColorData(), (size_t)W * H * 4);
		std::memcpy(depth_inout, fb.getDepthData(), (size_t)W * H * sizeof(float));
		if (seconds_out) *seconds_out = secs;
		// the cached pipeline must not keep pointers to this call's stack objects
		pipe->setCamera(nullptr); pipe->setFramebuffer(nullptr); pipe->setShader(nullptr);
		return 0;
	} catch (const std::exception& e) {
		fprintf(stderr, "axr_ref_render: %s\n", e.what());
		return -1;
	}
}

// Stage probe: Pipeline::clipTriangle on one triangle. in: 3 x (Vertex 14 f32 + clipPos 4 f32);
// out: up to 24 x 18 f32. Returns the number of output vertices (multiple of 3).
int axr_ref_clip_triangle(const float* in3x18, float* out24x18) {
	static_assert(sizeof(AR::ClippedVertex) == 72, "ClippedVertex layout");
	std::array<AR::ClippedVertex, AR::MAX_CLIPPED_VERTS> arr{};
	std::memcpy((void*)arr.data(), in3x18, 3 * sizeof(AR::ClippedVertex));
	size_t n = 3;
	AR::Pipeline::clipTriangle(n, arr);
	std::memcpy(out24x18, arr.data(), n * sizeof(AR::ClippedVertex));
	return (int)n;
}

// Stage probe: back-face decision + Triangle constructor outputs (screenPos[3], ndcZ[3], bbox[4]) = 13 f32.
int axr_ref_triangle_setup(const float* in3x18, int w, int h, float* out13) {
	AR::ClippedVertex v[3];
	std::memcpy((void*)v, in3x18, sizeof v);
	int back = AR::Triangle::isBackface(v[0], v[1], v[2], w, h) ? 1 : 0;
	AR::Triangle t(std::array<AR::ClippedVertex, 3>{v[0], v[1], v[2]}, w, h, nullptr);
	for (int i = 0; i < 3; ++i) { out13[i * 2] = t.screenPos[i].x; out13[i * 2 + 1] = t.screenPos[i].y; out13[6 + i] = t.ndcZ[i]; }
	out13[9] = t.minX; out13[10] = t.minY; out13[11] = t.maxX; out13[12] = t.maxY;
	return back;
}

// Stage probe: Texture::sample (nearest, V-flip) through a real AR::Texture.
int axr_ref_texture_sample(const uint8_t* rgba, int w, int h, const float* uv, int n, float* out_rgba) {
	auto t = make_texture(rgba, w, h, 9);
	if (!t) return -1;
	for (int i = 0; i < n; ++i) {
		glm::vec4 c = t->sample(glm::vec2(uv[i * 2], uv[i * 2 + 1]));
		out_rgba[i * 4 + 0] = c.x; out_rgba[i * 4 + 1] = c.y; out_rgba[i * 4 + 2] = c.z; out_rgba[i * 4 + 3] = c.w;
	}
	return 0;
}

// OBJ/MTL ingestion through the reference's own loader (src/mesh.cpp): returns the vertex/index
// arrays (incl. its generated tangents/bitangents) the hot path consumes.
void* axr_ref_mesh_load(const char* path) {
	try { return new AR::Mesh(std::string(path)); } catch (const std::exception& e) {
		fprintf(stderr, "axr_ref_mesh_load: %s\n", e.what());
		return nullptr;
	}
}
void axr_ref_mesh_counts(void* h, uint64_t* n_verts, uint64_t* n_faces) {
	auto* m = (AR::Mesh*)h;
	*n_verts = m->getVertices().size();
	*n_faces = m->getFaces().size();
}
void axr_ref_mesh_copy(void* h, float* vertices, uint32_t* indices) {
	auto* m = (AR::Mesh*)h;
	std::memcpy(vertices, m->getVertices().data(), m->getVertices().size() * sizeof(AR::Vertex));
	size_t k = 0;
	for (const auto& f : m->getFaces()) for (int i = 0; i < 3; ++i) indices[k++] = f.vertexIndices[i];
}
void axr_ref_mesh_free(void* h) { delete (AR::Mesh*)h; }

}  // extern "C"
