// IShader device functors: the shipped AxiomR shaders (reference include/shaders/shaders.hpp) restated as
// structs with the same two entry points as the plugin contract (reference include/IShader.hpp:30-46):
//   VertexOutput vertex(const Vertex&, int)                      -> Shader::vertex(uniforms, pos, attr, VsOut&)
//   bool fragment(vec3& bar, vec4& color, const VSTransformedTriangle&) -> Shader::fragment(...), true = discard
// The raster/shade kernel is templated on the functor, so a further IShader subclass is added by writing one more
// struct here and one more case in the dispatch of axr_kernels.cu.
#pragma once
#include "axr_raster.cuh"

namespace axr {

struct TexRef { const uchar4* data; int w, h; };

// Material of one group (reference include/mesh.hpp:20-34): diffuse, bump, metallic, roughness, ao + Ns
struct Material {
	TexRef tex[5];
	float specular_exponent;
	int pad;
};

struct Uniforms {
	m4 mvp;          // viewProj * model (reference src/tiled_pipeline.cpp:149)
	m4 model;
	m3 normal_mat;   // mat3(transpose(inverse(model))) — FlatShader::vertex :36-37, a pure function of `model`
	v3 cam_pos;
	v3 light_dir;
	v3 light_color;
	int sampler;     // 0 nearest (reference), 1 bilinear (extension)
};

// VertexOutput (reference include/IShader.hpp:11-17) as a flat array of varyings (zNDC is never read by the pipeline).
// Every shipped fragment shader starts by combining the three VertexOutputs with the barycentrics as
// bar.x*v0 + bar.y*v1 + bar.z*v2 (component-wise for vec2/vec3/mat3), i.e. ((v0*al) + (v1*be)) + v2*ga per float.
// The kernel performs exactly that combination (same order, same rounding) vertex by vertex, so only one VertexOutput
// is live at a time, and hands the interpolated varyings to Shader::fragment.
//   FlatShader : [0..2] normal
//   Phong / PBR: [0..1] uv, [2..4] worldPos, [5..13] tbn columns T,B,N
constexpr int VARY_FLAT = 3;
constexpr int VARY_TBN = 14;
constexpr int VARY_CUTOUT = 5;  // CutoutShader: [0..1] uv, [2..4] normal

// ------------------------------------------------------------------ Texture::sample (reference include/texture.hpp:12-34)
__device__ __forceinline__ v4 texel(const TexRef& t, int x, int y) {
	uchar4 p = __ldg(t.data + (size_t)y * (size_t)t.w + (size_t)x);
	const float inv255 = 1.0f / 255.0f;
	return V4(p.x * inv255, p.y * inv255, p.z * inv255, p.w * inv255);
}
__device__ __forceinline__ v4 sample_nearest(const TexRef& t, float u, float v) {
	if (!t.data) return V4(0, 0, 0, 1);
	int x = cvtt(u * (float)(t.w - 1));
	int y = cvtt(v * (float)(t.h - 1));
	x = max(0, min(x, t.w - 1));
	y = max(0, min(y, t.h - 1));
	y = t.h - 1 - y;
	return texel(t, x, y);
}
// EXTENSION without a reference counterpart: clamp-to-edge bilinear, lerp x then y with glm::mix order
__device__ __forceinline__ v4 sample_bilinear(const TexRef& t, float u, float v) {
	if (!t.data) return V4(0, 0, 0, 1);
	float wm = (float)(t.w - 1), hm = (float)(t.h - 1);
	float fx = u * wm, fy = v * hm;
	fx = fx > 0.0f ? fx : 0.0f; fx = fx < wm ? fx : wm;
	fy = fy > 0.0f ? fy : 0.0f; fy = fy < hm ? fy : hm;
	int x0 = (int)fx, y0 = (int)fy;
	int x1 = min(x0 + 1, t.w - 1), y1 = min(y0 + 1, t.h - 1);
	float tx = fx - (float)x0, ty = fy - (float)y0;
	v4 c00 = texel(t, x0, t.h - 1 - y0), c10 = texel(t, x1, t.h - 1 - y0);
	v4 c01 = texel(t, x0, t.h - 1 - y1), c11 = texel(t, x1, t.h - 1 - y1);
	return mix(mix(c00, c10, tx), mix(c01, c11, tx), ty);
}
// The sampler mode is a compile-time parameter of the shading kernel: no per-sample branch, and the four texel fetches of
// every bilinear tap of every map of a fragment are independent loads the scheduler can issue together.
template <int SMP>
__device__ __forceinline__ v4 sample(const TexRef& t, float u, float v) {
	return SMP ? sample_bilinear(t, u, v) : sample_nearest(t, u, v);
}

__device__ __forceinline__ v3 xyz(v4 v) { return V3(v.x, v.y, v.z); }

// PhongShader::vertex :147-168 == PBRShader::vertex :259-282
__device__ __forceinline__ void vertex_tbn(const Uniforms& u, v3 pos, v3 n, v3 t, v3 b, float uvx, float uvy, float* o) {
	o[0] = uvx; o[1] = uvy;
	v3 w = xyz(mul(u.model, V4(pos.x, pos.y, pos.z, 1.0f)));
	v3 T = normalize(xyz(mul(u.model, V4(t.x, t.y, t.z, 0.0f))));
	v3 B = normalize(xyz(mul(u.model, V4(b.x, b.y, b.z, 0.0f))));
	v3 N = normalize(xyz(mul(u.model, V4(n.x, n.y, n.z, 0.0f))));
	o[2] = w.x; o[3] = w.y; o[4] = w.z;
	o[5] = T.x; o[6] = T.y; o[7] = T.z;
	o[8] = B.x; o[9] = B.y; o[10] = B.z;
	o[11] = N.x; o[12] = N.y; o[13] = N.z;
}

// --------------------------------------------------------------------------------------------- FlatShader :19-61
struct FlatShader {
	static constexpr int NV = VARY_FLAT;
	static constexpr bool DISCARDS = false;  // fragment() never returns true: visibility does not depend on shading
	__device__ __forceinline__ static void vertex(const Uniforms& u, v3 pos, v3 n, v3 t, v3 b, float uvx, float uvy, float* o) {
		v3 r = mul(u.normal_mat, n);
		o[0] = r.x; o[1] = r.y; o[2] = r.z;
	}
	template <int SMP>
	__device__ __forceinline__ static bool fragment(const Uniforms& u, const Material& m, const float* var, v4& color) {
		v3 n = normalize(V3(var[0], var[1], var[2]));
		float intensity = clampf(dot(-u.light_dir, n), 0.0f, 1.0f);
		float c = 1.0f * intensity;
		color = V4(c, c, c, c);
		return false;
	}
};

// --------------------------------------------------------------------------------------------- PhongShader :136-250
struct PhongShader {
	static constexpr int NV = VARY_TBN;
	static constexpr bool DISCARDS = false;
	__device__ __forceinline__ static void vertex(const Uniforms& u, v3 pos, v3 n, v3 t, v3 b, float uvx, float uvy, float* o) {
		vertex_tbn(u, pos, n, t, b, uvx, uvy, o);
	}
	template <int SMP>
	__device__ __forceinline__ static bool fragment(const Uniforms& u, const Material& m, const float* var, v4& color) {
		const float uvx = var[0], uvy = var[1];
		v4 nm = sample<SMP>(m.tex[1], uvx, uvy);
		v4 albedo = sample<SMP>(m.tex[0], uvx, uvy);
		v3 nms = normalize(xyz(nm) * 2.0f - V3(1.0f, 1.0f, 1.0f));
		v3 T = V3(var[5], var[6], var[7]);
		v3 N = normalize(V3(var[11], var[12], var[13]));
		v3 Tn = normalize(T - N * dot(N, T));
		v3 Bn = cross(N, Tn);
		m3 ftbn;
		ftbn.c[0] = Tn; ftbn.c[1] = Bn; ftbn.c[2] = N;
		v3 normal = normalize(mul(ftbn, nms));
		v3 fragPos = V3(var[2], var[3], var[4]);
		v3 viewDir = normalize(u.cam_pos - fragPos);
		v3 lightDir = -u.light_dir;
		v3 ambient = u.light_color * 0.1f;
		float diff = maxf(dot(normal, lightDir), 0.0f);
		v3 diffuse = u.light_color * diff;
		v3 I = -lightDir;
		v3 reflectDir = I - normal * dot(normal, I) * 2.0f;  // glm::reflect
		float spec = powf(maxf(dot(viewDir, reflectDir), 0.0f), m.specular_exponent * 50.0f);
		v3 specular = u.light_color * (0.5f * spec);
		v3 fc = ((ambient + diffuse) + specular) * xyz(albedo);
		color = V4(fc.x, fc.y, fc.z, 1.0f);
		return false;
	}
};

// --------------------------------------------------------------------------------------------- PBRShader :252-423
struct PBRShader {
	static constexpr int NV = VARY_TBN;
	static constexpr bool DISCARDS = false;
	__device__ __forceinline__ static void vertex(const Uniforms& u, v3 pos, v3 n, v3 t, v3 b, float uvx, float uvy, float* o) {
		vertex_tbn(u, pos, n, t, b, uvx, uvy, o);
	}
	template <int SMP>
	__device__ __forceinline__ static bool fragment(const Uniforms& u, const Material& m, const float* var, v4& color) {
		const float PI = 3.14159265358979323846264338327950288f;
		const float uvx = var[0], uvy = var[1];
		v4 nm = sample<SMP>(m.tex[1], uvx, uvy);
		v4 al4 = sample<SMP>(m.tex[0], uvx, uvy);
		float metallic = sample<SMP>(m.tex[2], uvx, uvy).x;
		float roughness = sample<SMP>(m.tex[3], uvx, uvy).x;
		float ao = sample<SMP>(m.tex[4], uvx, uvy).x;
		v3 nms = normalize(V3(nm.x * 2.0f - 1.0f, nm.y * 2.0f - 1.0f, nm.z * 2.0f - 1.0f));
		m3 tbn;
		tbn.c[0] = V3(var[5], var[6], var[7]); tbn.c[1] = V3(var[8], var[9], var[10]); tbn.c[2] = V3(var[11], var[12], var[13]);
		v3 albedo = V3(powf(al4.x, 2.2f), powf(al4.y, 2.2f), powf(al4.z, 2.2f));
		float roughness2 = roughness * roughness;
		float roughness4 = roughness2 * roughness2;
		float oneMinusMetallic = 1.0f - metallic;
		v3 normal = normalize(mul(tbn, nms));  // :345 uses the raw interpolated TBN; the re-orthogonalised basis (:316-323) is dead
		v3 fragPos = V3(var[2], var[3], var[4]);
		v3 viewDir = normalize(u.cam_pos - fragPos);
		v3 lightDir = -u.light_dir;
		v3 halfwayDir = normalize(lightDir + viewDir);
		float NdotH = maxf(dot(normal, halfwayDir), 0.0f);
		float NdotV = maxf(dot(normal, viewDir), 0.0f);
		float NdotL = maxf(dot(normal, lightDir), 0.0f);
		v3 c04 = V3(0.04f, 0.04f, 0.04f);
		v3 F0 = c04 + (albedo - c04) * metallic;
		float NdotH2 = NdotH * NdotH;
		float denomPart = (NdotH2 * (roughness4 - 1.0f) + 1.0f);
		float NDF = roughness4 / (PI * denomPart * denomPart);
		float r = roughness + 1.0f;
		float k = (r * r) / 8.0f;
		float NdotV_k = NdotV * (1.0f - k) + k;
		float NdotL_k = NdotL * (1.0f - k) + k;
		float G = (NdotV / NdotV_k) * (NdotL / NdotL_k);
		float om = 1.0f - maxf(dot(halfwayDir, normal), 0.0f);
		float term = om * om;
		term *= term;
		term *= om;
		v3 one = V3(1.0f, 1.0f, 1.0f);
		v3 F = F0 + (one - F0) * term;
		v3 numerator = F * NDF * G;
		float denom = 4.0f * NdotV * NdotL + 0.0001f;
		v3 specular = numerator / denom;
		v3 kD = (one - F) * oneMinusMetallic;
		v3 diffuse = kD * albedo * (1.0f / PI);
		v3 ambient = V3(0.03f, 0.03f, 0.03f) * albedo * ao;
		v3 fc = ambient + (diffuse + specular) * u.light_color * NdotL;
		fc = fc / (fc + one);
		const float g = 1.0f / 2.2f;
		color = V4(powf(fc.x, g), powf(fc.y, g), powf(fc.z, g), 1.0f);
		return false;
	}
};

// --------------------------------------------------------------------------------------------- CutoutShader
// Not a reference shader: the reference ships none whose fragment() returns true, so its discard branch
// (src/tiled_pipeline.cpp:571-577) is exercised with an IShader of our own, defined against the reference's plugin
// contract in oracle/ref_harness.cpp (struct CutoutShader) and restated here: FlatShader's vertex stage plus uv, Lambert
// times the diffuse texel, fragments whose texel alpha is below 0.5 are discarded. DISCARDS = true makes the draw take the
// depth-peeling path (k_setup_raster<true> + the peel branch of k_tile_shade).
struct CutoutShader {
	static constexpr int NV = VARY_CUTOUT;
	static constexpr bool DISCARDS = true;
	__device__ __forceinline__ static void vertex(const Uniforms& u, v3 pos, v3 n, v3 t, v3 b, float uvx, float uvy, float* o) {
		v3 r = mul(u.normal_mat, n);
		o[0] = uvx; o[1] = uvy;
		o[2] = r.x; o[3] = r.y; o[4] = r.z;
	}
	template <int SMP>
	__device__ __forceinline__ static bool fragment(const Uniforms& u, const Material& m, const float* var, v4& color) {
		v4 texl = sample<SMP>(m.tex[0], var[0], var[1]);
		if (texl.w < 0.5f) return true;
		v3 n = normalize(V3(var[2], var[3], var[4]));
		float intensity = clampf(dot(-u.light_dir, n), 0.0f, 1.0f);
		color = V4(texl.x * intensity, texl.y * intensity, texl.z * intensity, 1.0f);
		return false;
	}
};

}  // namespace axr
