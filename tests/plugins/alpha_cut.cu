// Test plug-in 2: a shader whose fragment() discards (returns true) — the same arithmetic as the library's CutoutShader, which
// oracle/ref_harness.cpp runs through the UNMODIFIED reference pipeline as an IShader subclass; supplied here from outside the
// library, so the frames must be identical to AXR_SHADER_CUTOUT's. The alpha threshold comes from Uniforms::user[0].
#include "axr_shader_plugin.cuh"

struct AlphaCut {
	static constexpr int NV = 5;
	static constexpr bool DISCARDS = true;
	static constexpr bool HAS_FAST = false;
	static constexpr unsigned TEXTURES = 1u;
	__device__ __forceinline__ static void vertex(const axr::Uniforms& u, axr::v3 pos, axr::v3 n, axr::v3 t, axr::v3 b, float uvx, float uvy, float* o) {
		const axr::v3 r = axr::mul(u.normal_mat, n);
		o[0] = uvx; o[1] = uvy;
		o[2] = r.x; o[3] = r.y; o[4] = r.z;
	}
	template <int SMP>
	__device__ __forceinline__ static bool fragment(const axr::Uniforms& u, const axr::Material& m, const float* var, axr::v4& color) {
		const axr::v4 texl = axr::sample<SMP>(m.tex[0], var[0], var[1]);
		if (texl.w < u.user[0]) return true;
		const axr::v3 n = axr::normalize(axr::V3(var[2], var[3], var[4]));
		const float intensity = axr::clampf(axr::dot(-u.light_dir, n), 0.0f, 1.0f);
		color = axr::V4(texl.x * intensity, texl.y * intensity, texl.z * intensity, 1.0f);
		return false;
	}
};
AXR_SHADER_PLUGIN(AlphaCut)
