// IShader device functors: the shipped AxiomR shaders (reference include/shaders/shaders.hpp) restated as
// structs with the same two entry points as the plugin contract (reference include/IShader.hpp:30-46):
//   VertexOutput vertex(const Vertex&, int)                      -> Shader::vertex(uniforms, pos, attr, VsOut&)
//   bool fragment(vec3& bar, vec4& color, const VSTransformedTriangle&) -> Shader::fragment(...), true = discard
// The raster/shade kernel is templated on the functor, so a further IShader subclass is added by writing one more
// struct here and one more case in the dispatch of axr_kernels.cu.
#pragma once
#include "axr_raster.cuh"

namespace axr {

// Texels live in HBM in a tiled order (AXR_TEX_TILED, default): a 128 B line holds an 8 x 4 texel block, each of its four 32 B
// sectors a 4 x 2 block, so the taps of a bilinear lookup and the lookups of neighbouring pixels fall into few lines / sectors
// (linear rows: ~15 distinct lines per warp-wide tap gather on the 4K scene; tiled: about half). tiles_x = ceil(w / 8).
#ifndef AXR_TEX_TILED
#define AXR_TEX_TILED 1
#endif
struct TexRef { const uchar4* data; int w, h; int tiles_x; };
// element offset of texel (x, y); y counts texture rows (row 0 = image top)
__host__ __device__ __forceinline__ unsigned tex_offset_x(int x) {
#if AXR_TEX_TILED
	return (((unsigned)x >> 3) << 5) | ((unsigned)x & 3u) | (((unsigned)x & 4u) << 1);
#else
	return (unsigned)x;
#endif
}
__host__ __device__ __forceinline__ unsigned tex_offset_y(int y, int w, int tiles_x) {
#if AXR_TEX_TILED
	return ((unsigned)y >> 2) * ((unsigned)tiles_x << 5) + ((((unsigned)y & 1u) << 2) | (((unsigned)y & 2u) << 3));
#else
	return (unsigned)y * (unsigned)w;
#endif
}

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
	float user[8];   // axr_set_shader_user: free parameters of a plug-in shader (axr_shader_plugin.cuh)
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
__device__ __forceinline__ unsigned texel_word(const TexRef& t, int x, int y) {
	return __ldg(reinterpret_cast<const unsigned*>(t.data) + (tex_offset_y(y, t.w, t.tiles_x) + tex_offset_x(x)));
}
__device__ __forceinline__ v4 unpack_texel(unsigned p) {
	const float inv255 = 1.0f / 255.0f;
	return V4(u8_to_f32(p, 0) * inv255, u8_to_f32(p, 1) * inv255, u8_to_f32(p, 2) * inv255, u8_to_f32(p, 3) * inv255);
}
__device__ __forceinline__ v4 texel(const TexRef& t, int x, int y) { return unpack_texel(texel_word(t, x, y)); }
__device__ __forceinline__ void nearest_xy(const TexRef& t, float u, float v, int& x, int& y) {
	x = cvtt(u * (float)(t.w - 1));
	y = cvtt(v * (float)(t.h - 1));
	x = max(0, min(x, t.w - 1));
	y = max(0, min(y, t.h - 1));
	y = t.h - 1 - y;
}
__device__ __forceinline__ v4 sample_nearest(const TexRef& t, float u, float v) {
	if (!t.data) return V4(0, 0, 0, 1);
	int x, y;
	nearest_xy(t, u, v, x, y);
	return texel(t, x, y);
}
// EXTENSION without a reference counterpart: clamp-to-edge bilinear, lerp x then y with glm::mix order.
// The four taps and the two weights of a lookup (exact in both colour modes: they select texels).
struct BilinearTaps { int x0, x1, r0, r1; float tx, ty; };  // r0 / r1: texture rows (already flipped) of the y0 / y1 taps
__device__ __forceinline__ BilinearTaps bilinear_taps(const TexRef& t, float u, float v) {
	BilinearTaps b;
	float wm = (float)(t.w - 1), hm = (float)(t.h - 1);
	float fx = u * wm, fy = v * hm;
	fx = fx > 0.0f ? fx : 0.0f; fx = fx < wm ? fx : wm;
	fy = fy > 0.0f ? fy : 0.0f; fy = fy < hm ? fy : hm;
	int y0;
	float flx, fly;
	floor_small(fx, b.x0, flx);  // == (int)fx, (float)x0: 0 <= fx < 2^22
	floor_small(fy, y0, fly);
	b.x1 = min(b.x0 + 1, t.w - 1);
	const int y1 = min(y0 + 1, t.h - 1);
	b.tx = fx - flx; b.ty = fy - fly;
	b.r0 = t.h - 1 - y0; b.r1 = t.h - 1 - y1;
	return b;
}
__device__ __forceinline__ v4 sample_bilinear(const TexRef& t, float u, float v) {
	if (!t.data) return V4(0, 0, 0, 1);
	const BilinearTaps b = bilinear_taps(t, u, v);
	v4 c00 = texel(t, b.x0, b.r0), c10 = texel(t, b.x1, b.r0);
	v4 c01 = texel(t, b.x0, b.r1), c11 = texel(t, b.x1, b.r1);
	return mix(mix(c00, c10, b.tx), mix(c01, c11, b.tx), b.ty);
}
// The sampler mode is a compile-time parameter of the shading kernel: no per-sample branch, and the four texel fetches of
// every bilinear tap of every map of a fragment are independent loads the scheduler can issue together.
template <int SMP>
__device__ __forceinline__ v4 sample(const TexRef& t, float u, float v) {
	return SMP ? sample_bilinear(t, u, v) : sample_nearest(t, u, v);
}

// ---- fast colour mode (see axr_math.cuh, namespace fm): same texels, same weights; the blend runs on the raw bytes with fused
// multiply-adds and is scaled by 1/255 once. A byte k of a texel word is moved into the mantissa of 2^15 (one PRMT, value
// 32768 + b, spacing 2^-8), so the three lerps need no unpacking subtraction; the offset leaves with the final scale.
struct TexLookup { unsigned i00, i10, i01, i11; float tx, ty; };  // element offsets of the taps + weights (nearest: i00 only)
template <int SMP>
__device__ __forceinline__ TexLookup make_lookup(const TexRef& t, float u, float v) {
	TexLookup l;
	if (SMP) {
		const BilinearTaps b = bilinear_taps(t, u, v);
		const unsigned ox0 = tex_offset_x(b.x0), ox1 = tex_offset_x(b.x1);
		const unsigned oy0 = tex_offset_y(b.r0, t.w, t.tiles_x), oy1 = tex_offset_y(b.r1, t.w, t.tiles_x);
		l.i00 = oy0 + ox0; l.i10 = oy0 + ox1; l.i01 = oy1 + ox0; l.i11 = oy1 + ox1;
		l.tx = b.tx; l.ty = b.ty;
	} else {
		int x, y;
		nearest_xy(t, u, v, x, y);
		l.i00 = l.i10 = l.i01 = l.i11 = tex_offset_y(y, t.w, t.tiles_x) + tex_offset_x(x);
		l.tx = l.ty = 0.0f;
	}
	return l;
}
struct RawTexels { unsigned p00, p10, p01, p11; float tx, ty; };
template <int SMP>
__device__ __forceinline__ RawTexels fetch_texels(const TexRef& t, const TexLookup& l) {
	const unsigned* d = reinterpret_cast<const unsigned*>(t.data);
	RawTexels r;
	r.p00 = __ldg(d + l.i00);
	if (SMP) { r.p10 = __ldg(d + l.i10); r.p01 = __ldg(d + l.i01); r.p11 = __ldg(d + l.i11); }
	else r.p10 = r.p01 = r.p11 = r.p00;
	r.tx = l.tx; r.ty = l.ty;
	return r;
}
__device__ __forceinline__ float byte_at_2p15(unsigned word, int k) {
#ifdef __CUDA_ARCH__
	return __uint_as_float(__byte_perm(word, 0x47000000u, 0x7404u + ((unsigned)k << 4)));
#else
	return 32768.0f + (float)((word >> (8 * k)) & 0xffu);
#endif
}
template <int SMP>
__device__ __forceinline__ float blend_channel(const RawTexels& r, int k) {
	const float inv255 = 1.0f / 255.0f;
	if (SMP) {
		const float a = byte_at_2p15(r.p00, k), b = byte_at_2p15(r.p10, k), c = byte_at_2p15(r.p01, k), d = byte_at_2p15(r.p11, k);
		const float top = fm::lerp(a, b, r.tx), bot = fm::lerp(c, d, r.tx);
		return fm::fma(fm::lerp(top, bot, r.ty), inv255, -32768.0f * inv255);
	}
	return u8_to_f32(r.p00, k) * inv255;
}
template <int SMP>
__device__ __forceinline__ v3 blend_rgb(const RawTexels& r) { return V3(blend_channel<SMP>(r, 0), blend_channel<SMP>(r, 1), blend_channel<SMP>(r, 2)); }

__device__ __forceinline__ v3 xyz(v4 v) { return V3(v.x, v.y, v.z); }

// One vertex as the shading stage sees it (AR::Vertex, reference include/mesh.hpp:9-18)
struct VIn { v3 pos, n, t, b; float u, v; };

// ---- fast colour mode: Shader::shade_fast() replaces vertex() x3 + the barycentric combination + fragment() for one pixel.
// Only the uv interpolation keeps the reference's individually rounded form (it selects texels); the rest uses namespace fm
// and the linearity of the vertex stage: sum_k w_k (M p_k) == M (sum_k w_k p_k), evaluated once per pixel instead of per vertex.
__device__ __forceinline__ void interp_uv(const float w[3], const VIn v[3], float& uvx, float& uvy) {
	uvx = ((v[0].u * w[0]) + (v[1].u * w[1])) + v[2].u * w[2];
	uvy = ((v[0].v * w[0]) + (v[1].v * w[1])) + v[2].v * w[2];
}
__device__ __forceinline__ v3 interp_fast(const float w[3], v3 a, v3 b, v3 c) { return fm::madd(fm::madd(fm::scale(a, w[0]), b, w[1]), c, w[2]); }
// sum_k w_k normalize(M3 d_k): the vertex stage normalises each transformed direction before the combination
__device__ __forceinline__ v3 interp_unit_fast(const m4& model, const float w[3], v3 a, v3 b, v3 c) {
	const v3 ma = fm::mul3(model, a), mb = fm::mul3(model, b), mc = fm::mul3(model, c);
	const float sa = fm::rsq(fm::dotf(ma, ma)) * w[0], sb = fm::rsq(fm::dotf(mb, mb)) * w[1], sc = fm::rsq(fm::dotf(mc, mc)) * w[2];
	return fm::madd(fm::madd(fm::scale(ma, sa), mb, sb), mc, sc);
}

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
	static constexpr bool HAS_FAST = true;
	template <int SMP>
	__device__ __forceinline__ static bool shade_fast(const Uniforms& u, const Material& m, const float w[3], const VIn v[3], v4& color) {
		const v3 n = fm::nrm(fm::mul3(u.normal_mat, interp_fast(w, v[0].n, v[1].n, v[2].n)));
		const float c = clampf(fm::dotf(-u.light_dir, n), 0.0f, 1.0f);
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
	static constexpr bool HAS_FAST = true;
	template <int SMP>
	__device__ __forceinline__ static bool shade_fast(const Uniforms& u, const Material& m, const float w[3], const VIn v[3], v4& color) {
		float uvx, uvy;
		interp_uv(w, v, uvx, uvy);
		// both maps of a material usually have the same size: one set of taps and weights then serves both
		const TexLookup l1 = make_lookup<SMP>(m.tex[1], uvx, uvy);
		TexLookup l0 = l1;
		if (m.tex[0].w != m.tex[1].w || m.tex[0].h != m.tex[1].h) l0 = make_lookup<SMP>(m.tex[0], uvx, uvy);
		const RawTexels rn = fetch_texels<SMP>(m.tex[1], l1), ra = fetch_texels<SMP>(m.tex[0], l0);
		const v3 fragPos = fm::affine(u.model, interp_fast(w, v[0].pos, v[1].pos, v[2].pos));
		const v3 T = interp_unit_fast(u.model, w, v[0].t, v[1].t, v[2].t);
		const v3 N = fm::nrm(interp_unit_fast(u.model, w, v[0].n, v[1].n, v[2].n));
		const v3 Tn = fm::nrm(fm::madd(T, N, -fm::dotf(N, T)));
		const v3 Bn = fm::crs(N, Tn);
		const v3 nm = blend_rgb<SMP>(rn);
		const v3 nms = fm::nrm(V3(fm::fma(nm.x, 2.0f, -1.0f), fm::fma(nm.y, 2.0f, -1.0f), fm::fma(nm.z, 2.0f, -1.0f)));
		const v3 normal = fm::nrm(fm::madd(fm::madd(fm::scale(Tn, nms.x), Bn, nms.y), N, nms.z));
		const v3 viewDir = fm::nrm(u.cam_pos - fragPos);
		const v3 lightDir = -u.light_dir;
		const float ndl = fm::dotf(normal, lightDir);
		const float diff = maxf(ndl, 0.0f);
		const v3 reflectDir = fm::madd(u.light_dir, normal, 2.0f * ndl);  // reflect(-L, n) = -L - n * dot(n, -L) * 2
		const float sd = maxf(fm::dotf(viewDir, reflectDir), 0.0f);
		const float e = m.specular_exponent * 50.0f;
		// 2^(e log2 x) carries the absolute error of the SFU log2 times e: fine for the exponents of matte materials, not for the
		// thousands of a mirror-like Ns; those keep powf (a per-material, hence warp-uniform, branch). powf(x, 0) == 1 for every x.
		const float spec = (e == 0.0f) ? 1.0f : ((fabsf(e) <= 64.0f) ? fm::pow_pos(sd, e) : powf(sd, e));
		const float k = fm::fma(0.5f, spec, 0.1f + diff);  // ambient 0.1 + diffuse + specularStrength 0.5 * spec, times lightColor
		const v3 albedo = blend_rgb<SMP>(ra);
		color = V4(u.light_color.x * k * albedo.x, u.light_color.y * k * albedo.y, u.light_color.z * k * albedo.z, 1.0f);
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
	static constexpr bool HAS_FAST = true;
	template <int SMP>
	__device__ __forceinline__ static bool shade_fast(const Uniforms& u, const Material& m, const float w[3], const VIn v[3], v4& color) {
		const float PI = 3.14159265358979323846264338327950288f;
		float uvx, uvy;
		interp_uv(w, v, uvx, uvy);
		TexLookup l[5];
		l[1] = make_lookup<SMP>(m.tex[1], uvx, uvy);
#pragma unroll
		for (int i = 0; i < 5; ++i)
			if (i != 1) {
				l[i] = l[1];
				if (m.tex[i].w != m.tex[1].w || m.tex[i].h != m.tex[1].h) l[i] = make_lookup<SMP>(m.tex[i], uvx, uvy);
			}
		const RawTexels rn = fetch_texels<SMP>(m.tex[1], l[1]), ra = fetch_texels<SMP>(m.tex[0], l[0]);
		const RawTexels rm = fetch_texels<SMP>(m.tex[2], l[2]), rr = fetch_texels<SMP>(m.tex[3], l[3]), ro = fetch_texels<SMP>(m.tex[4], l[4]);
		const v3 fragPos = fm::affine(u.model, interp_fast(w, v[0].pos, v[1].pos, v[2].pos));
		const v3 T = interp_unit_fast(u.model, w, v[0].t, v[1].t, v[2].t);
		const v3 B = interp_unit_fast(u.model, w, v[0].b, v[1].b, v[2].b);
		const v3 N = interp_unit_fast(u.model, w, v[0].n, v[1].n, v[2].n);
		const v3 nm = blend_rgb<SMP>(rn);
		const v3 nms = fm::nrm(V3(fm::fma(nm.x, 2.0f, -1.0f), fm::fma(nm.y, 2.0f, -1.0f), fm::fma(nm.z, 2.0f, -1.0f)));
		const v3 normal = fm::nrm(fm::madd(fm::madd(fm::scale(T, nms.x), B, nms.y), N, nms.z));
		const v3 al = blend_rgb<SMP>(ra);
		const v3 albedo = V3(fm::pow_pos(al.x, 2.2f), fm::pow_pos(al.y, 2.2f), fm::pow_pos(al.z, 2.2f));
		const float metallic = blend_channel<SMP>(rm, 0), roughness = blend_channel<SMP>(rr, 0), ao = blend_channel<SMP>(ro, 0);
		const float roughness2 = roughness * roughness, roughness4 = roughness2 * roughness2;
		const v3 viewDir = fm::nrm(u.cam_pos - fragPos);
		const v3 lightDir = -u.light_dir;
		const v3 halfwayDir = fm::nrm(lightDir + viewDir);
		const float NdotH = maxf(fm::dotf(normal, halfwayDir), 0.0f);
		const float NdotV = maxf(fm::dotf(normal, viewDir), 0.0f);
		const float NdotL = maxf(fm::dotf(normal, lightDir), 0.0f);
		const v3 F0 = V3(fm::fma(albedo.x - 0.04f, metallic, 0.04f), fm::fma(albedo.y - 0.04f, metallic, 0.04f),
		                 fm::fma(albedo.z - 0.04f, metallic, 0.04f));
		const float denomPart = fm::fma(NdotH * NdotH, roughness4 - 1.0f, 1.0f);
		const float NDF = roughness4 * fm::rcp(PI * denomPart * denomPart);
		const float r = roughness + 1.0f;
		const float k = (r * r) * 0.125f;
		const float NdotV_k = fm::fma(NdotV, 1.0f - k, k), NdotL_k = fm::fma(NdotL, 1.0f - k, k);
		const float G = (NdotV * NdotL) * fm::rcp(NdotV_k * NdotL_k);
		const float om = 1.0f - NdotH;
		const float om2 = om * om;
		const float term = om2 * om2 * om;
		const v3 F = V3(fm::fma(1.0f - F0.x, term, F0.x), fm::fma(1.0f - F0.y, term, F0.y), fm::fma(1.0f - F0.z, term, F0.z));
		const float sg = NDF * G * fm::rcp(fm::fma(4.0f * NdotV, NdotL, 0.0001f));  // specular = F * sg
		const float oneMinusMetallic = 1.0f - metallic;
		const float invPI = 1.0f / PI;
		v3 fc;
		{
			const float dx = (1.0f - F.x) * oneMinusMetallic * albedo.x * invPI, dy = (1.0f - F.y) * oneMinusMetallic * albedo.y * invPI,
			            dz = (1.0f - F.z) * oneMinusMetallic * albedo.z * invPI;
			const float ax = 0.03f * albedo.x * ao, ay = 0.03f * albedo.y * ao, az = 0.03f * albedo.z * ao;
			fc = V3(fm::fma(fm::fma(F.x, sg, dx) * u.light_color.x, NdotL, ax), fm::fma(fm::fma(F.y, sg, dy) * u.light_color.y, NdotL, ay),
			        fm::fma(fm::fma(F.z, sg, dz) * u.light_color.z, NdotL, az));
		}
		fc = V3(fc.x * fm::rcp(fc.x + 1.0f), fc.y * fm::rcp(fc.y + 1.0f), fc.z * fm::rcp(fc.z + 1.0f));
		const float g = 1.0f / 2.2f;
		color = V4(fm::pow_pos(fc.x, g), fm::pow_pos(fc.y, g), fm::pow_pos(fc.z, g), 1.0f);
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
	static constexpr bool HAS_FAST = false;  // its discard decides coverage: exact arithmetic only
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
