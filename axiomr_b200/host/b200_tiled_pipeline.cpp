// See b200_tiled_pipeline.hpp. Replaces reference src/tiled_pipeline.cpp:143-322 with calls into libaxr_b200.so.
#include "b200_tiled_pipeline.hpp"

#include <cmath>
#include <stdexcept>
#include <string>

#include "axr_b200.h"
#include "camera.hpp"
#include "framebuffer.hpp"
#include "shaders/shaders.hpp"

namespace AR {

B200TiledPipeline::B200TiledPipeline(size_t /*threadsAvailable*/, Camera* cam, Framebuffer* fb, int cudaDevice)
	: Pipeline(cam, fb), m_Device(cudaDevice) {}

B200TiledPipeline::~B200TiledPipeline() {
	if (m_Ctx) axr_destroy(m_Ctx);
}

void B200TiledPipeline::fail(const char* what, int code) {
	std::string msg = std::string(what) + " failed (" + std::to_string(code) + "): " + axr_last_error(m_Ctx);
	throw std::runtime_error(msg);
}

void B200TiledPipeline::ensureContext() {
	const int w = (int)m_Framebuffer->getWidth(), h = (int)m_Framebuffer->getHeight();
	if (m_Ctx && w == m_CtxW && h == m_CtxH) return;
	if (m_Ctx) { axr_destroy(m_Ctx); m_Ctx = nullptr; m_Meshes.clear(); m_Textures.clear(); }
	axr_config cfg{};
	cfg.device = m_Device;
	cfg.width = w;
	cfg.height = h;
	cfg.sampler = AXR_SAMPLER_NEAREST;  // Texture::sample as shipped (reference include/texture.hpp:12-34)
	int rc = axr_create(&cfg, &m_Ctx);
	if (rc != AXR_OK) fail("axr_create", rc);
	axr_set_overlap(m_Ctx, 1);
	m_CtxW = w;
	m_CtxH = h;
}

int B200TiledPipeline::textureHandle(const Texture* tex) {
	if (!tex || tex->getWidth() <= 0 || tex->getHeight() <= 0) return AXR_NO_TEXTURE;
	auto it = m_Textures.find(tex);
	if (it != m_Textures.end()) return it->second;
	// Texture keeps its RGBA8 pixels private; getPixelRGBA (include/texture.hpp:35-45) returns byte * (1/255), which rounds back exactly.
	const int w = tex->getWidth(), h = tex->getHeight();
	std::vector<uint8_t> rgba((size_t)w * h * 4);
	for (int y = 0; y < h; ++y)
		for (int x = 0; x < w; ++x) {
			glm::vec4 p = tex->getPixelRGBA(x, y);
			uint8_t* o = &rgba[((size_t)y * w + x) * 4];
			o[0] = (uint8_t)std::lround(p.x * 255.0f);
			o[1] = (uint8_t)std::lround(p.y * 255.0f);
			o[2] = (uint8_t)std::lround(p.z * 255.0f);
			o[3] = (uint8_t)std::lround(p.w * 255.0f);
		}
	axr_tex t = AXR_NO_TEXTURE;
	int rc = axr_upload_texture(m_Ctx, rgba.data(), w, h, &t);
	if (rc != AXR_OK) fail("axr_upload_texture", rc);
	m_LastH2D += rgba.size();
	m_Textures.emplace(tex, t);
	return t;
}

int B200TiledPipeline::meshHandle(const Mesh& mesh) {
	auto it = m_Meshes.find(&mesh);
	if (it != m_Meshes.end()) return it->second;
	const auto& verts = mesh.getVertices();
	const auto& faces = mesh.getFaces();
	static_assert(sizeof(Vertex) == 56, "AR::Vertex layout (include/mesh.hpp:9-18)");
	std::vector<uint32_t> idx;
	idx.reserve(faces.size() * 3);
	for (const Face& f : faces) {
		if (f.vertexIndices.size() != 3) throw std::runtime_error("B200TiledPipeline: faces must be triangles (reference asserts this, src/tiled_pipeline.cpp:202)");
		idx.insert(idx.end(), f.vertexIndices.begin(), f.vertexIndices.end());
	}
	const auto& groups = mesh.getMaterialGroups();
	std::vector<axr_group> g;
	for (const MaterialGroup& mg : groups) g.push_back(axr_group{(uint64_t)mg.startIndex, (uint64_t)mg.faceCount});
	axr_mesh h = -1;
	int rc = axr_upload_mesh(m_Ctx, reinterpret_cast<const float*>(verts.data()), verts.size(), idx.data(), faces.size(),
	                         g.empty() ? nullptr : g.data(), (uint32_t)g.size(), &h);
	if (rc != AXR_OK) fail("axr_upload_mesh", rc);
	m_LastH2D += verts.size() * sizeof(Vertex) + idx.size() * sizeof(uint32_t);
	for (size_t i = 0; i < groups.size(); ++i) {
		const Material* m = mesh.getMaterial(groups[i].materialName);  // throws std::out_of_range like the reference (:179)
		rc = axr_set_material(m_Ctx, h, (uint32_t)i, textureHandle(m->diffuseTexture.get()), textureHandle(m->bumpTexture.get()),
		                      textureHandle(m->metallicTexture.get()), textureHandle(m->roughnessTexture.get()),
		                      textureHandle(m->aoTexture.get()), m->specularExponent);
		if (rc != AXR_OK) fail("axr_set_material", rc);
	}
	m_Meshes.emplace(&mesh, h);
	return h;
}

void B200TiledPipeline::invalidate(const Mesh& mesh) {
	auto it = m_Meshes.find(&mesh);
	if (it == m_Meshes.end()) return;
	if (m_Ctx) axr_free_mesh(m_Ctx, it->second);
	m_Meshes.erase(it);
}

void B200TiledPipeline::drawMesh(const glm::mat4& modelMatrix, const Mesh& mesh) {
	if (!m_Shader || !m_Camera || !m_Framebuffer) return;  // reference src/tiled_pipeline.cpp:146
	ensureContext();
	m_LastH2D = m_LastD2H = 0;

	// IShader subclass -> device functor + its public parameters (include/shaders/shaders.hpp:59-60,243-245,401-403)
	axr_shader_params sp{};
	int kind;
	if (auto* s = dynamic_cast<FlatShader*>(m_Shader)) {
		kind = AXR_SHADER_FLAT;
		sp.light_dir[0] = s->lightDirection.x; sp.light_dir[1] = s->lightDirection.y; sp.light_dir[2] = s->lightDirection.z;
		sp.light_color[0] = sp.light_color[1] = sp.light_color[2] = 1.0f;
	} else if (auto* p = dynamic_cast<PhongShader*>(m_Shader)) {
		kind = AXR_SHADER_PHONG;
		sp.light_dir[0] = p->lightDirection.x; sp.light_dir[1] = p->lightDirection.y; sp.light_dir[2] = p->lightDirection.z;
		sp.light_color[0] = p->lightColor.x; sp.light_color[1] = p->lightColor.y; sp.light_color[2] = p->lightColor.z;
	} else if (auto* b = dynamic_cast<PBRShader*>(m_Shader)) {
		kind = AXR_SHADER_PBR;
		sp.light_dir[0] = b->lightDirection.x; sp.light_dir[1] = b->lightDirection.y; sp.light_dir[2] = b->lightDirection.z;
		sp.light_color[0] = b->lightColor.x; sp.light_color[1] = b->lightColor.y; sp.light_color[2] = b->lightColor.z;
	} else if (auto* u = dynamic_cast<B200PluginShader*>(m_Shader)) {
		int rcp = axr_load_shader_plugin(m_Ctx, u->path.c_str(), &kind);  // opened once per context, then looked up by path
		if (rcp != AXR_OK) fail("axr_load_shader_plugin", rcp);
		sp.light_dir[0] = u->lightDirection.x; sp.light_dir[1] = u->lightDirection.y; sp.light_dir[2] = u->lightDirection.z;
		sp.light_color[0] = u->lightColor.x; sp.light_color[1] = u->lightColor.y; sp.light_color[2] = u->lightColor.z;
		rcp = axr_set_shader_user(m_Ctx, u->user, 8);
		if (rcp != AXR_OK) fail("axr_set_shader_user", rcp);
	} else {
		throw std::runtime_error("B200TiledPipeline: this IShader subclass has no device functor (no CPU fallback by design); "
		                         "compile it as a plug-in (include/axr_shader_plugin.cuh) and pass a B200PluginShader");
	}
	int rc = axr_set_shader(m_Ctx, kind, &sp, sizeof sp);
	if (rc != AXR_OK) fail("axr_set_shader", rc);

	const glm::mat4& viewProj = m_Camera->getViewProjectionMatrix();
	const glm::mat4 viewport = m_Camera->getViewportMatrix();
	const glm::vec3 camPos = m_Camera->getPosition();
	const float cp[3] = {camPos.x, camPos.y, camPos.z};
	rc = axr_set_uniforms(m_Ctx, &viewProj[0][0], &viewport[0][0], cp);  // glm::mat4 is 16 contiguous column-major floats
	if (rc != AXR_OK) fail("axr_set_uniforms", rc);

	const int h = meshHandle(mesh);
	const size_t npx = (size_t)m_Framebuffer->getWidth() * m_Framebuffer->getHeight();
	// nothing is uploaded: host depth of the visible pixels read through the zero-copy mapping (4 B each), passing pixels stored
	// the same way (8 B each); complete on return
	rc = axr_draw_mesh_host(m_Ctx, h, &modelMatrix[0][0], m_Framebuffer->getColorData(), m_Framebuffer->getDepthData());
	if (rc != AXR_OK) fail("axr_draw_mesh_host", rc);
	(void)npx;
	m_LastH2D += 16 * 4 * 3 + 12;
	m_LastD2H += 0;  // 8 bytes per updated pixel, written by the kernel
}

}  // namespace AR
