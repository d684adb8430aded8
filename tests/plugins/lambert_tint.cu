// Test plug-in 1: a shader the library does not ship. FlatShader's vertex stage plus uv; the fragment is Lambert times the diffuse
// texel times a tint taken from Uniforms::user. With tint = (1,1,1) and no discard it equals CutoutShader on an opaque texture, which
// is how tests/test_gpu_parity.py pins it (CutoutShader itself is pinned against the reference's IShader contract).
#include "axr_shader_plugin.cuh"

struct LambertTint {
	static constexpr int NV = 5;
	static constexpr bool DISCARDS = false;
	static constexpr bool HAS_FAST = false;
	static constexpr unsigned TEXTURES = 1u;  // diffuse
	__device__ __forceinline__ static void vertex(const axr::Uniforms& u, axr::v3 pos, axr::v3 n, axr::v3 t, axr::v3 b, float uvx, float uvy, float* o) {
		const axr::v3 r = axr::mul(u.normal_mat, n);
		o[0] = uvx; o[1] = uvy;
		o[2] = r.x; o[3] = r.y; o[4] = r.z;
	}
	template <int SMP>
	__device__ __forceinline__ static bool fragment(const axr::Uniforms& u, const axr::Material& m, const float* var, axr::v4& color) {
		const axr::v4 texl = axr::sample<SMP>(m.tex[0], var[0], var[1]);
		const axr::v3 n = axr::normalize(axr::V3(var[2], var[3], var[4]));
		const float intensity = axr::clampf(axr::dot(-u.light_dir, n), 0.0f, 1.0f);
		color = axr::V4(texl.x * intensity * u.user[0], texl.y * intensity * u.user[1], texl.z * intensity * u.user[2], 1.0f);
		return false;
	}
};
AXR_SHADER_PLUGIN(LambertTint)
