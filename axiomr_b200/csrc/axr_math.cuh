// Binary32 vector/matrix helpers in glm's evaluation order (host + device).
//
// Every function restates the published scalar glm definition the reference's hot path relies on
// (reference include/math.hpp:9-14 pulls glm with default packed types => scalar code), with each
// + and * individually rounded. The translation units that include this header are compiled with
// -fmad=false (device) and -ffp-contract=off (host): no FMA contraction, so results are bit-identical
// to the CPU reference built the same way.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define AXR_HD __host__ __device__ __forceinline__
#define AXR_D __device__ __forceinline__

namespace axr {

struct v3 { float x, y, z; };
struct v4 { float x, y, z, w; };
struct m3 { v3 c[3]; };  // column-major
struct m4 { v4 c[4]; };

AXR_HD v3 V3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
AXR_HD v4 V4(float x, float y, float z, float w) { v4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
AXR_HD v3 operator+(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
AXR_HD v3 operator-(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
AXR_HD v3 operator-(v3 a) { return V3(-a.x, -a.y, -a.z); }
AXR_HD v3 operator*(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
AXR_HD v3 operator*(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
AXR_HD v3 operator/(v3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
AXR_HD v3 operator/(v3 a, v3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
AXR_HD v4 operator+(v4 a, v4 b) { return V4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
AXR_HD v4 operator*(v4 a, float s) { return V4(a.x * s, a.y * s, a.z * s, a.w * s); }

AXR_HD float dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }  // compute_dot<vec3>
AXR_HD v3 normalize(v3 v) { return v * (1.0f / sqrtf(dot(v, v))); }          // v * inversesqrt(dot(v,v))
AXR_HD v3 cross(v3 x, v3 y) { return V3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
AXR_HD float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }
AXR_HD v3 mix(v3 x, v3 y, float a) { return x * (1.0f - a) + y * a; }
AXR_HD v4 mix(v4 x, v4 y, float a) { return x * (1.0f - a) + y * a; }

// type_mat4x4.inl operator*(mat4, vec4): (m0*x + m1*y) + (m2*z + m3*w)
AXR_HD v4 mul(const m4& m, v4 v) { return (m.c[0] * v.x + m.c[1] * v.y) + (m.c[2] * v.z + m.c[3] * v.w); }
// type_mat3x3.inl operator*(mat3, vec3): row sums left to right
AXR_HD v3 mul(const m3& m, v3 v) {
	return V3(m.c[0].x * v.x + m.c[1].x * v.y + m.c[2].x * v.z,
	          m.c[0].y * v.x + m.c[1].y * v.y + m.c[2].y * v.z,
	          m.c[0].z * v.x + m.c[1].z * v.y + m.c[2].z * v.z);
}
// type_mat4x4.inl operator*(mat4, mat4): ((a0*b.x + a1*b.y) + a2*b.z) + a3*b.w per column
AXR_HD m4 mul(const m4& a, const m4& b) {
	m4 r;
	for (int i = 0; i < 4; ++i) {
		v4 bc = b.c[i];
		r.c[i] = ((a.c[0] * bc.x + a.c[1] * bc.y) + a.c[2] * bc.z) + a.c[3] * bc.w;
	}
	return r;
}
AXR_HD m4 load_m4(const float* p) {
	m4 m;
	for (int c = 0; c < 4; ++c) m.c[c] = V4(p[c * 4], p[c * 4 + 1], p[c * 4 + 2], p[c * 4 + 3]);
	return m;
}
// func_matrix.inl compute_inverse<4,4,float>: cofactors, determinant from the first column
inline __host__ m4 inverse(const m4& mm) {
	float m[4][4];
	for (int c = 0; c < 4; ++c) { m[c][0] = mm.c[c].x; m[c][1] = mm.c[c].y; m[c][2] = mm.c[c].z; m[c][3] = mm.c[c].w; }
	float c00 = m[2][2] * m[3][3] - m[3][2] * m[2][3], c02 = m[1][2] * m[3][3] - m[3][2] * m[1][3], c03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
	float c04 = m[2][1] * m[3][3] - m[3][1] * m[2][3], c06 = m[1][1] * m[3][3] - m[3][1] * m[1][3], c07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
	float c08 = m[2][1] * m[3][2] - m[3][1] * m[2][2], c10 = m[1][1] * m[3][2] - m[3][1] * m[1][2], c11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
	float c12 = m[2][0] * m[3][3] - m[3][0] * m[2][3], c14 = m[1][0] * m[3][3] - m[3][0] * m[1][3], c15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
	float c16 = m[2][0] * m[3][2] - m[3][0] * m[2][2], c18 = m[1][0] * m[3][2] - m[3][0] * m[1][2], c19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
	float c20 = m[2][0] * m[3][1] - m[3][0] * m[2][1], c22 = m[1][0] * m[3][1] - m[3][0] * m[1][1], c23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
	const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
	const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
	const float a0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]}, a1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
	const float a2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]}, a3[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
	const float sa[4] = {+1, -1, +1, -1}, sb[4] = {-1, +1, -1, +1};
	float inv[4][4];
	for (int i = 0; i < 4; ++i) {
		float i0 = a1[i] * f0[i] - a2[i] * f1[i] + a3[i] * f2[i];
		float i1 = a0[i] * f0[i] - a2[i] * f3[i] + a3[i] * f4[i];
		float i2 = a0[i] * f1[i] - a1[i] * f3[i] + a3[i] * f5[i];
		float i3 = a0[i] * f2[i] - a1[i] * f4[i] + a2[i] * f5[i];
		inv[0][i] = i0 * sa[i]; inv[1][i] = i1 * sb[i]; inv[2][i] = i2 * sa[i]; inv[3][i] = i3 * sb[i];
	}
	float d0 = m[0][0] * inv[0][0], d1 = m[0][1] * inv[1][0], d2 = m[0][2] * inv[2][0], d3 = m[0][3] * inv[3][0];
	float ood = 1.0f / ((d0 + d1) + (d2 + d3));
	m4 r;
	for (int c = 0; c < 4; ++c) r.c[c] = V4(inv[c][0] * ood, inv[c][1] * ood, inv[c][2] * ood, inv[c][3] * ood);
	return r;
}

// float -> int the way the reference's x86-64 build converts (cvttss2si): truncation, INT_MIN for NaN / out of range.
AXR_HD int cvtt(float f) { return (f > -2147483904.0f && f < 2147483648.0f) ? (int)f : (-2147483647 - 1); }
AXR_HD float clampf(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }  // std::clamp
AXR_HD float maxf(float a, float b) { return (a < b) ? b : a; }                                    // std::max
AXR_HD float min3f(float a, float b, float c) { float m = a; if (b < m) m = b; if (c < m) m = c; return m; }  // std::min({..})
AXR_HD float max3f(float a, float b, float c) { float m = a; if (m < b) m = b; if (m < c) m = c; return m; }  // std::max({..})

// ------------------------------------------------------------------ exact integer <-> float moves off the conversion pipe
// I2F / F2I issue at an eighth of the FP32 rate on sm_100; for small non-negative integers the same results come from the
// FADD pipe: 2^23 + n has n in its low mantissa bits. Bit-identical to the casts they replace (within the stated ranges).
// byte k of a packed RGBA8 word -> float, exact (replaces I2F.U8)
AXR_D float u8_to_f32(unsigned word, int k) {
#ifdef __CUDA_ARCH__
	return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7650u + (unsigned)k)) - 8388608.0f;
#else
	return (float)((word >> (8 * k)) & 0xffu);
#endif
}
// floor of 0 <= f < 2^22 as int and as float: (int)f and (float)(int)f without F2I / I2F (f + 2^23 rounded toward zero)
AXR_D void floor_small(float f, int& i, float& fl) {
#ifdef __CUDA_ARCH__
	const float t = __fadd_rz(f, 8388608.0f);
	i = (int)(__float_as_uint(t) & 0x7fffffu);
	fl = t - 8388608.0f;
#else
	i = (int)f;
	fl = (float)i;
#endif
}
// non-negative int < 2^23 -> float, exact (replaces I2F)
AXR_D float small_int_to_f32(int i) {
#ifdef __CUDA_ARCH__
	return __uint_as_float(0x4B000000u + (unsigned)i) - 8388608.0f;
#else
	return (float)i;
#endif
}

// ------------------------------------------------------------------ fast colour math
// The parity contract (BASELINE.json north_star) is exact coverage and depth, and 8-bit colour within 1 LSB. Everything that
// feeds coverage, depth or the choice of a texel stays in the individually-rounded forms above; arithmetic that only feeds
// the colour of a pixel may use these: fused multiply-adds and the SFU approximations (rel. error ~1e-7 .. 2^-22).
namespace fm {
AXR_D float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
AXR_D float rsq(float x) {
#ifdef __CUDA_ARCH__
	float r;
	asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
#else
	return 1.0f / sqrtf(x);
#endif
}
AXR_D float rcp(float x) {
#ifdef __CUDA_ARCH__
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
#else
	return 1.0f / x;
#endif
}
// x^y for x >= 0 (or NaN), small |y|: 2^(y log2 x). y == 0 is handled by the callers (powf(x, 0) == 1 for every x).
AXR_D float pow_pos(float x, float y) {
#ifdef __CUDA_ARCH__
	float l, r;
	asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
	asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l * y));
	return r;
#else
	return powf(x, y);
#endif
}
AXR_D float dotf(v3 a, v3 b) { return fma(a.z, b.z, fma(a.y, b.y, a.x * b.x)); }
AXR_D v3 scale(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
AXR_D v3 madd(v3 acc, v3 a, float s) { return V3(fma(a.x, s, acc.x), fma(a.y, s, acc.y), fma(a.z, s, acc.z)); }  // acc + a*s
AXR_D v3 nrm(v3 a) { return scale(a, rsq(dotf(a, a))); }
AXR_D v3 crs(v3 x, v3 y) { return V3(fma(x.y, y.z, -(y.y * x.z)), fma(x.z, y.x, -(y.z * x.x)), fma(x.x, y.y, -(y.x * x.y))); }
// upper-left 3x3 of a column-major mat4 times a direction
AXR_D v3 mul3(const m4& m, v3 v) {
	return V3(fma(m.c[2].x, v.z, fma(m.c[1].x, v.y, m.c[0].x * v.x)), fma(m.c[2].y, v.z, fma(m.c[1].y, v.y, m.c[0].y * v.x)),
	          fma(m.c[2].z, v.z, fma(m.c[1].z, v.y, m.c[0].z * v.x)));
}
AXR_D v3 mul3(const m3& m, v3 v) {
	return V3(fma(m.c[2].x, v.z, fma(m.c[1].x, v.y, m.c[0].x * v.x)), fma(m.c[2].y, v.z, fma(m.c[1].y, v.y, m.c[0].y * v.x)),
	          fma(m.c[2].z, v.z, fma(m.c[1].z, v.y, m.c[0].z * v.x)));
}
// model * (p, 1), xyz only
AXR_D v3 affine(const m4& m, v3 p) {
	return V3(fma(m.c[2].x, p.z, fma(m.c[1].x, p.y, fma(m.c[0].x, p.x, m.c[3].x))), fma(m.c[2].y, p.z, fma(m.c[1].y, p.y, fma(m.c[0].y, p.x, m.c[3].y))),
	          fma(m.c[2].z, p.z, fma(m.c[1].z, p.y, fma(m.c[0].z, p.x, m.c[3].z))));
}
AXR_D float lerp(float a, float b, float t) { return fma(t, b - a, a); }
}  // namespace fm

}  // namespace axr
