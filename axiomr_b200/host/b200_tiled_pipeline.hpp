// AR::B200TiledPipeline — drop-in replacement for AR::TiledPipeline (reference include/tiled_pipeline.hpp:58-62) that runs
// drawMesh on a B200 through the C ABI of include/axr_b200.h.
//
// It compiles INSIDE the reference's tree (against its own pipeline.hpp / IShader.hpp / mesh.hpp / camera.hpp /
// framebuffer.hpp / shaders/shaders.hpp and glm) and keeps the reference's constructor and virtual interface:
//     TiledPipeline(size_t threadsAvailable, Camera*, Framebuffer*)   ->   B200TiledPipeline(size_t, Camera*, Framebuffer*)
//     void drawMesh(const glm::mat4& modelMatrix, const Mesh& mesh) override
// so `m_Pipeline = std::make_unique<TiledPipeline>(...)` at reference src/renderer.cpp:71 becomes
// `std::make_unique<B200TiledPipeline>(...)` (or `using TiledPipeline = B200TiledPipeline;`), see INTEGRATION.md.
//
// Semantics kept: borrowed pointers, silent return without shader/camera/framebuffer (src/tiled_pipeline.cpp:146), composites
// onto the framebuffer's current contents with the strict depth test, complete on return. Differences: the thread count is
// ignored; host virtuals cannot run on the device and there is deliberately no CPU fallback, so an IShader subclass other than
// Flat/Phong/PBRShader is given as a B200PluginShader (below): the handle of a device functor its author compiled with
// tools/build_shader_plugin.py; any other subclass throws std::runtime_error; capacity problems surface as std::runtime_error with the axr error text instead
// of std::bad_alloc from the 16 MB arena.
#pragma once
#include <cstddef>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "pipeline.hpp"

struct axr_ctx;

namespace AR {

// Pipeline::setShader(IShader*) for a shader of the user's own: the device code lives in a plug-in library (include/axr_shader_plugin.cuh),
// this object carries its path and parameters. Its host virtuals are never called by B200TiledPipeline (and throw if the reference's
// CPU pipelines are handed one).
struct B200PluginShader : public IShader {
	explicit B200PluginShader(const std::string& pluginPath) : path(pluginPath) {}
	std::string path;
	glm::vec3 lightDirection{0.0f, -1.0f, 0.0f};
	glm::vec3 lightColor{1.0f, 1.0f, 1.0f};
	float user[8] = {};  // Uniforms::user of the functor
private:
	VertexOutput vertex(const Vertex&, int) override { throw std::runtime_error("B200PluginShader has device code only"); }
	bool fragment(glm::vec3&, glm::vec4&, const VSTransformedTriangle&) override { throw std::runtime_error("B200PluginShader has device code only"); }
};

class B200TiledPipeline : public Pipeline {
public:
	B200TiledPipeline(size_t threadsAvailable, Camera* cam, Framebuffer* fb, int cudaDevice = 0);
	~B200TiledPipeline();
	B200TiledPipeline(const B200TiledPipeline&) = delete;
	B200TiledPipeline& operator=(const B200TiledPipeline&) = delete;
	void drawMesh(const glm::mat4& modelMatrix, const Mesh& mesh) override;
	// The reference reads the host Mesh on every drawMesh; this adapter uploads it on first use and keeps the device copy, keyed by
	// the Mesh's address. After editing a Mesh in place (or destroying it and building another at the same address), call
	// invalidate(): the device copy is dropped and the next drawMesh uploads the current contents.
	void invalidate(const Mesh& mesh);

	// bytes moved host<->device by the last drawMesh (framebuffer round trip + first-use mesh / texture uploads)
	size_t lastH2DBytes() const { return m_LastH2D; }
	size_t lastD2HBytes() const { return m_LastD2H; }

private:
	void ensureContext();
	int meshHandle(const Mesh& mesh);
	int textureHandle(const Texture* tex);
	[[noreturn]] void fail(const char* what, int code);

	axr_ctx* m_Ctx = nullptr;
	int m_Device = 0;
	int m_CtxW = 0, m_CtxH = 0;
	std::unordered_map<const Mesh*, int> m_Meshes;        // uploaded once, keyed by identity (the reference re-reads the host Mesh every call)
	std::unordered_map<const Texture*, int> m_Textures;
	size_t m_LastH2D = 0, m_LastD2H = 0;
};

}  // namespace AR
