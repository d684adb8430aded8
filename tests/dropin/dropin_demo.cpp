// TEST PROGRAM (not product): renders the same OBJ/MTL scene, loaded by the reference's own Mesh/Texture loaders, once with the
// reference's AR::TiledPipeline (CPU) and once with AR::B200TiledPipeline (CUDA, through the C ABI) in ONE process, through the
// reference's own public API, and compares the two host framebuffers. Built by `make -C oracle dropin` into oracle/_ref/.
//   usage: dropin_demo <scene.obj> <width> <height> <shader 0|1|2>          FlatShader | PhongShader | PBRShader
//          dropin_demo <scene.obj> <width> <height> 3 <lambert_tint.so>   a shader of the user's own: the reference runs the IShader
//                                                                          subclass below, the B200 pipeline the plug-in functor
//                                                                          tests/plugins/lambert_tint.cu compiled from the same formulas
#include "prelude.hpp"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "tiled_pipeline.hpp"
#include "shaders/shaders.hpp"
#include "camera.hpp"
#include "framebuffer.hpp"
#include "mesh.hpp"
#include "b200_tiled_pipeline.hpp"

// A user's IShader written against the reference's plugin contract (include/IShader.hpp:30-46): FlatShader's normal handling plus the
// diffuse texel and a tint. Its device twin is tests/plugins/lambert_tint.cu.
struct LambertTintShader : public AR::IShader {
	glm::vec3 lightDirection{0.0f, -1.0f, 0.0f};
	glm::vec3 tint{1.0f, 1.0f, 1.0f};
	AR::VertexOutput vertex(const AR::Vertex& v, int) override {
		AR::VertexOutput o;
		o.uv = v.uv;
		o.normal = glm::mat3(glm::transpose(glm::inverse(model))) * v.normal;
		return o;
	}
	bool fragment(glm::vec3& bar, glm::vec4& color, const AR::VSTransformedTriangle& tri) override {
		glm::vec2 uv = bar.x * tri[0].uv + bar.y * tri[1].uv + bar.z * tri[2].uv;
		glm::vec4 texel = material->diffuseTexture->sample(uv);
		glm::vec3 n = bar.x * tri[0].normal + bar.y * tri[1].normal + bar.z * tri[2].normal;
		n = glm::normalize(n);
		float intensity = std::clamp(glm::dot(-lightDirection, n), 0.0f, 1.0f);
		color = glm::vec4(texel.x * intensity * tint.x, texel.y * intensity * tint.y, texel.z * intensity * tint.z, 1.0f);
		return false;
	}
};

namespace AR {
bool Window::isMouseButtonDown(int) const { return false; }
bool Window::isKeyDown(char) const { return false; }
}

int main(int argc, char** argv) {
	if (argc < 5) { fprintf(stderr, "usage: %s scene.obj W H shader\n", argv[0]); return 2; }
	const int W = atoi(argv[2]), H = atoi(argv[3]), kind = atoi(argv[4]);
	try {
		AR::Mesh mesh{std::string(argv[1])};
		AR::Camera cam(glm::vec3(0, 0, 5), glm::vec3(0, 0, 0), 60.0f, (float)W / (float)H);  // reference src/renderer.cpp:68-69
		cam.setViewport(0, 0, W, H);
		cam.update(0.0f);
		AR::FlatShader flat; AR::PhongShader phong; AR::PBRShader pbr;
		glm::vec3 L = glm::normalize(glm::vec3(-0.3f, -1.0f, -0.5f)), LC(0.6f, 0.6f, 0.6f);
		flat.lightDirection = L; phong.lightDirection = L; phong.lightColor = LC; pbr.lightDirection = L; pbr.lightColor = LC;
		AR::IShader* sh = kind == 0 ? (AR::IShader*)&flat : kind == 1 ? (AR::IShader*)&phong : (AR::IShader*)&pbr;
		// kind 3: the same user shader twice — host virtuals for the reference pipeline, a device plug-in for the B200 one
		LambertTintShader userCpu;
		userCpu.lightDirection = L; userCpu.tint = glm::vec3(0.9f, 0.55f, 0.3f);
		AR::B200PluginShader userGpu(argc > 5 ? argv[5] : "");
		userGpu.lightDirection = L; userGpu.user[0] = userCpu.tint.x; userGpu.user[1] = userCpu.tint.y; userGpu.user[2] = userCpu.tint.z;
		AR::IShader* shGpu = sh;
		if (kind == 3) {
			if (argc < 6) { fprintf(stderr, "shader 3 needs the plug-in library\n"); return 2; }
			sh = &userCpu; shGpu = &userGpu;
		}
		glm::mat4 model = glm::rotate(glm::mat4(1.0f), 0.5f, glm::vec3(0, 1, 0));

		AR::Framebuffer fbRef(W, H, true), fbGpu(W, H, true);
		for (AR::Framebuffer* fb : {&fbRef, &fbGpu}) { fb->clearColor({0, 0, 0, 255}); fb->clearDepth(); }

		std::unique_ptr<AR::TiledPipeline> ref(new AR::TiledPipeline(4, &cam, &fbRef));
		ref->setShader(sh);
		auto t0 = std::chrono::steady_clock::now();
		ref->drawMesh(model, mesh);
		double msRef = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

		AR::Pipeline* gpu = new AR::B200TiledPipeline(4, &cam, &fbGpu);  // used through the base-class interface, like Renderer does
		gpu->setShader(shGpu);
		gpu->drawMesh(model, mesh);  // first call uploads + caches the mesh
		for (AR::Framebuffer* fb : {&fbGpu}) { fb->clearColor({0, 0, 0, 255}); fb->clearDepth(); }
		t0 = std::chrono::steady_clock::now();
		gpu->drawMesh(model, mesh);
		double msGpu = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		delete gpu;

		size_t n = (size_t)W * H, cov = 0, covMis = 0, depthBits = 0, col1 = 0, colBad = 0;
		const float *dr = fbRef.getDepthData(), *dg = fbGpu.getDepthData();
		const uint8_t *cr = fbRef.getColorData(), *cg = fbGpu.getColorData();
		for (size_t i = 0; i < n; ++i) {
			bool a = std::isfinite(dr[i]), b = std::isfinite(dg[i]);
			cov += a;
			covMis += (a != b);
			depthBits += std::memcmp(&dr[i], &dg[i], 4) != 0;
			int worst = 0;
			for (int c = 0; c < 4; ++c) worst = std::max(worst, std::abs((int)cr[i * 4 + c] - (int)cg[i * 4 + c]));
			col1 += worst == 1;
			colBad += worst > 1;
		}
		printf("faces=%zu covered=%zu coverage_mismatch=%zu depth_bit_mismatch=%zu color_off_by_1=%zu color_off_by_more=%zu ref_ms=%.2f b200_ms=%.2f\n",
		       mesh.getFaces().size(), cov, covMis, depthBits, col1, colBad, msRef, msGpu);
		bool ok = cov > 0 && covMis == 0 && depthBits == 0 && colBad == 0 && col1 <= n / 1000;
		printf(ok ? "PARITY OK\n" : "PARITY FAILED\n");
		return ok ? 0 : 1;
	} catch (const std::exception& e) {
		fprintf(stderr, "dropin_demo: %s\n", e.what());
		return 3;
	}
}
