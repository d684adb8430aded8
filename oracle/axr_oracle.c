/* TEST INFRASTRUCTURE — not part of the product path; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this.
 *
 * Plain-C restatement of AxiomR's tiled rasterisation path (TiledPipeline::drawMesh and everything
 * it calls), following the reference function by function. Citations are relative to /root/reference.
 * All arithmetic is IEEE binary32 with every + and * individually rounded: build with
 * -ffp-contract=off (oracle/Makefile), matching the shimmed reference build in oracle/_ref.
 *
 * PARITY PINNING: the reference ships no tests, golden images or known-answer vectors
 * (tests/CMakeLists.txt is empty), so this restatement is pinned against OUTPUTS OF THE REFERENCE
 * ITSELF: tests/test_oracle_vs_ref.py compares it bit-for-bit with oracle/_ref (the unmodified
 * reference sources) wherever that library is present, and tests/golden/ holds fixtures generated
 * from oracle/_ref by tests/golden/make_golden.py. glm is an un-vendored, un-pinned dependency of the
 * reference (external/CMakeLists.txt:12-14); its published scalar definitions (0.9.9.x/1.0.x) are
 * restated inline below, operation order included.
 *
 * The `sampler == 1` (bilinear) mode is an EXTENSION with no reference counterpart:
 * it is compared against this file only and every report says so.
 */
#include "axr_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define TILE_SIZE 16          /* include/tiled_pipeline.hpp:28 */
#define MAX_CLIPPED_VERTS 24  /* include/pipeline.hpp:17 */

typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } v4;
typedef struct { v3 c[3]; } m3;   /* column-major, glm::mat3 */
typedef struct { v4 c[4]; } m4;   /* column-major, glm::mat4 */

/* include/mesh.hpp:9-18 */
typedef struct { float position[3], uv[2], normal[3], tangent[3], bitangent[3]; } vertex_t;
/* include/pipeline.hpp:13-16 */
typedef struct { vertex_t v; v4 clip; } cvert_t;
/* include/IShader.hpp:11-17 (zNDC is never read by the pipeline) */
typedef struct { v3 normal; float uv[2]; v3 world; m3 tbn; } vsout_t;
/* include/tiled_pipeline.hpp:14-26 */
typedef struct {
	float sx[3], sy[3], ndcz[3];
	float minx, miny, maxx, maxy;
	vsout_t vs[3];  /* IShader::vertex outputs; the reference recomputes them per (triangle, tile), src/tiled_pipeline.cpp:407-409 */
} tri_t;

/* ------------------------------------------------------------------ glm restatement */
static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 ld3(const float* p) { return V3(p[0], p[1], p[2]); }
static inline v3 add3(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub3(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul3(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 scl3(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 neg3(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline float dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }          /* compute_dot<vec3> */
static inline v3 norm3(v3 v) { return scl3(v, 1.0f / sqrtf(dot3(v, v))); }                    /* v * inversesqrt(dot) */
static inline v3 cross3(v3 x, v3 y) { return V3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
static inline v4 V4(float x, float y, float z, float w) { v4 r = {x, y, z, w}; return r; }
static inline v4 add4(v4 a, v4 b) { return V4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static inline v4 scl4(v4 a, float s) { return V4(a.x * s, a.y * s, a.z * s, a.w * s); }
static inline float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }        /* glm::mix */
static inline v3 mix3(v3 x, v3 y, float a) { return add3(scl3(x, 1.0f - a), scl3(y, a)); }
static inline v4 mix4(v4 x, v4 y, float a) { return add4(scl4(x, 1.0f - a), scl4(y, a)); }
/* type_mat4x4.inl operator*(mat4, vec4): (m0*x + m1*y) + (m2*z + m3*w) */
static inline v4 m4_mul_v4(const m4* m, v4 v) {
	return add4(add4(scl4(m->c[0], v.x), scl4(m->c[1], v.y)), add4(scl4(m->c[2], v.z), scl4(m->c[3], v.w)));
}
/* type_mat3x3.inl operator*(mat3, vec3): row sums left to right */
static inline v3 m3_mul_v3(const m3* m, v3 v) {
	return V3(m->c[0].x * v.x + m->c[1].x * v.y + m->c[2].x * v.z,
	          m->c[0].y * v.x + m->c[1].y * v.y + m->c[2].y * v.z,
	          m->c[0].z * v.x + m->c[1].z * v.y + m->c[2].z * v.z);
}
static inline m3 m3_scl(const m3* m, float s) { m3 r = {{scl3(m->c[0], s), scl3(m->c[1], s), scl3(m->c[2], s)}}; return r; }
static inline m3 m3_add(const m3* a, const m3* b) { m3 r = {{add3(a->c[0], b->c[0]), add3(a->c[1], b->c[1]), add3(a->c[2], b->c[2])}}; return r; }
static m4 ld_m4(const float* p) {
	m4 m;
	for (int c = 0; c < 4; ++c) m.c[c] = V4(p[c * 4], p[c * 4 + 1], p[c * 4 + 2], p[c * 4 + 3]);
	return m;
}
static void st_m4(const m4* m, float* p) {
	for (int c = 0; c < 4; ++c) { p[c * 4] = m->c[c].x; p[c * 4 + 1] = m->c[c].y; p[c * 4 + 2] = m->c[c].z; p[c * 4 + 3] = m->c[c].w; }
}
/* type_mat4x4.inl operator*(mat4, mat4): ((a0*b.x + a1*b.y) + a2*b.z) + a3*b.w per column */
static m4 m4_mul(const m4* a, const m4* b) {
	m4 r;
	for (int i = 0; i < 4; ++i) {
		v4 bc = b->c[i];
		r.c[i] = add4(add4(add4(scl4(a->c[0], bc.x), scl4(a->c[1], bc.y)), scl4(a->c[2], bc.z)), scl4(a->c[3], bc.w));
	}
	return r;
}
/* func_matrix.inl compute_inverse<4,4,float> */
static m4 m4_inverse(const m4* mm) {
	float m[4][4];
	for (int c = 0; c < 4; ++c) { m[c][0] = mm->c[c].x; m[c][1] = mm->c[c].y; m[c][2] = mm->c[c].z; m[c][3] = mm->c[c].w; }
	float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
	float Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
	float Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
	float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
	float Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
	float Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
	float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
	float Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
	float Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
	float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
	float Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
	float Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
	float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
	float Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
	float Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
	float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
	float Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
	float Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
	float Fac0[4] = {Coef00, Coef00, Coef02, Coef03}, Fac1[4] = {Coef04, Coef04, Coef06, Coef07};
	float Fac2[4] = {Coef08, Coef08, Coef10, Coef11}, Fac3[4] = {Coef12, Coef12, Coef14, Coef15};
	float Fac4[4] = {Coef16, Coef16, Coef18, Coef19}, Fac5[4] = {Coef20, Coef20, Coef22, Coef23};
	float Vec0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]}, Vec1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
	float Vec2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]}, Vec3[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
	static const float SignA[4] = {+1, -1, +1, -1}, SignB[4] = {-1, +1, -1, +1};
	float inv[4][4];
	for (int i = 0; i < 4; ++i) {
		float Inv0 = Vec1[i] * Fac0[i] - Vec2[i] * Fac1[i] + Vec3[i] * Fac2[i];
		float Inv1 = Vec0[i] * Fac0[i] - Vec2[i] * Fac3[i] + Vec3[i] * Fac4[i];
		float Inv2 = Vec0[i] * Fac1[i] - Vec1[i] * Fac3[i] + Vec3[i] * Fac5[i];
		float Inv3 = Vec0[i] * Fac2[i] - Vec1[i] * Fac4[i] + Vec2[i] * Fac5[i];
		inv[0][i] = Inv0 * SignA[i]; inv[1][i] = Inv1 * SignB[i]; inv[2][i] = Inv2 * SignA[i]; inv[3][i] = Inv3 * SignB[i];
	}
	float d0 = m[0][0] * inv[0][0], d1 = m[0][1] * inv[1][0], d2 = m[0][2] * inv[2][0], d3 = m[0][3] * inv[3][0];
	float Dot1 = (d0 + d1) + (d2 + d3);
	float OneOverDeterminant = 1.0f / Dot1;
	m4 r;
	for (int c = 0; c < 4; ++c)
		r.c[c] = V4(inv[c][0] * OneOverDeterminant, inv[c][1] * OneOverDeterminant, inv[c][2] * OneOverDeterminant, inv[c][3] * OneOverDeterminant);
	return r;
}

void axo_mat4_mul(const float a[16], const float b[16], float out[16]) {
	m4 A = ld_m4(a), B = ld_m4(b), R = m4_mul(&A, &B);
	st_m4(&R, out);
}
void axo_mat4_inverse(const float m[16], float out[16]) {
	m4 A = ld_m4(m), R = m4_inverse(&A);
	st_m4(&R, out);
}

/* float -> int as the reference's x86-64 build performs it (cvttss2si): truncation toward zero,
 * INT_MIN for NaN and out-of-range values. Spelled out so the restatement is defined C. */
static inline int cvtt(float f) {
	if (!(f > -2147483904.0f && f < 2147483648.0f)) return (-2147483647 - 1);
	return (int)f;
}
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline float fmin3(float a, float b, float c) { float m = a; if (b < m) m = b; if (c < m) m = c; return m; }  /* std::min({a,b,c}) */
static inline float fmax3(float a, float b, float c) { float m = a; if (m < b) m = b; if (m < c) m = c; return m; }  /* std::max({a,b,c}) */
static inline float clampf(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }              /* std::clamp */
static inline float maxf(float a, float b) { return (a < b) ? b : a; }                                               /* std::max */

/* ------------------------------------------------------------------ src/tiled_pipeline.cpp:15-23 */
static inline float clampW(float w) {
	const float tiny = 1e-6f;
	if (fabsf(w) < tiny) w = (w < 0.0f) ? -tiny : tiny;
	return w;
}

/* ------------------------------------------------------------------ clipping: src/pipeline.cpp */
/* :274-285 */
static inline float distFunc(v4 v, int plane) {
	switch (plane) {
	case 0: return v.x + v.w;
	case 1: return v.w - v.x;
	case 2: return v.y + v.w;
	case 3: return v.w - v.y;
	case 4: return v.z + v.w;
	case 5: return v.w - v.z;
	default: return 0.0f;
	}
}
/* :243-272 */
static cvert_t interpolateVertices(const cvert_t* a, const cvert_t* b, float t) {
	cvert_t o;
	v3 p = mix3(ld3(a->v.position), ld3(b->v.position), t);
	v3 n = norm3(mix3(ld3(a->v.normal), ld3(b->v.normal), t));
	v3 tg = norm3(mix3(ld3(a->v.tangent), ld3(b->v.tangent), t));
	v3 bt = norm3(mix3(ld3(a->v.bitangent), ld3(b->v.bitangent), t));
	o.v.position[0] = p.x; o.v.position[1] = p.y; o.v.position[2] = p.z;
	o.v.normal[0] = n.x; o.v.normal[1] = n.y; o.v.normal[2] = n.z;
	o.v.tangent[0] = tg.x; o.v.tangent[1] = tg.y; o.v.tangent[2] = tg.z;
	o.v.bitangent[0] = bt.x; o.v.bitangent[1] = bt.y; o.v.bitangent[2] = bt.z;
	o.clip = mix4(a->clip, b->clip, t);
	float w0 = a->clip.w, w1 = b->clip.w;
	float u0x = a->v.uv[0] * w0, u0y = a->v.uv[1] * w0;
	float u1x = b->v.uv[0] * w1, u1y = b->v.uv[1] * w1;
	float ux = u0x * (1.0f - t) + u1x * t, uy = u0y * (1.0f - t) + u1y * t;
	float iw = mixf(w0, w1, t);
	o.v.uv[0] = ux / iw;
	o.v.uv[1] = uy / iw;
	return o;
}
/* :302-370 (McGuire 2011); v0..v2 are modified in place, v3 is the possible fourth vertex */
static int clipTriangleSinglePlane(int plane, cvert_t* v0, cvert_t* v1, cvert_t* v2, cvert_t* v3_) {
	float d0 = distFunc(v0->clip, plane), d1 = distFunc(v1->clip, plane), d2 = distFunc(v2->clip, plane);
	if (d0 < 0.f && d1 < 0.f && d2 < 0.f) return 0;
	if (d0 >= 0.f && d1 >= 0.f && d2 >= 0.f) { *v3_ = *v0; return 3; }
	cvert_t tv; float td;
#define SWAPV(a, b) do { tv = *(a); *(a) = *(b); *(b) = tv; } while (0)
#define SWAPD(a, b) do { td = (a); (a) = (b); (b) = td; } while (0)
	if (d1 >= 0.f && !(d0 >= 0.f)) {
		SWAPV(v0, v1); SWAPD(d0, d1);
		SWAPV(v1, v2); SWAPD(d1, d2);
	} else if (d2 >= 0.f && !(d1 >= 0.f)) {
		SWAPV(v2, v1); SWAPD(d2, d1);
		SWAPV(v1, v0); SWAPD(d1, d0);
	}
#undef SWAPV
#undef SWAPD
	float denom02 = d0 - d2;
	float t02 = (fabsf(denom02) < 1e-7f) ? 0.5f : (d0 / denom02);
	*v3_ = interpolateVertices(v0, v2, t02);
	if (d1 >= 0.f) {
		float denom12 = d1 - d2;
		float t12 = (fabsf(denom12) < 1e-7f) ? 0.5f : (d1 / denom12);
		*v2 = interpolateVertices(v1, v2, t12);
		return 4;
	}
	float denom01 = d0 - d1;
	float t01 = (fabsf(denom01) < 1e-7f) ? 0.5f : (d0 / denom01);
	*v1 = interpolateVertices(v0, v1, t01);
	*v2 = *v3_;
	return 3;
}
/* :176-228; `cur` holds n vertices on entry, the clipped list on exit; returns the new count */
static int clipTriangle(int n, cvert_t* cur, cvert_t* next) {
	cvert_t* a = cur; cvert_t* b = next;
	for (int plane = 0; plane < 6; ++plane) {
		int cnt = 0;
		for (int i = 0; i < n; i += 3) {
			cvert_t q;
			int r = clipTriangleSinglePlane(plane, &a[i], &a[i + 1], &a[i + 2], &q);
			if (r == 3) {
				if (cnt + 3 <= MAX_CLIPPED_VERTS) { b[cnt] = a[i]; b[cnt + 1] = a[i + 1]; b[cnt + 2] = a[i + 2]; cnt += 3; }
			} else if (r == 4) {
				if (cnt + 6 <= MAX_CLIPPED_VERTS) {
					b[cnt] = a[i]; b[cnt + 1] = a[i + 1]; b[cnt + 2] = a[i + 2];
					b[cnt + 3] = a[i]; b[cnt + 4] = a[i + 2]; b[cnt + 5] = q;
					cnt += 6;
				}
			}
		}
		n = cnt;
		cvert_t* t = a; a = b; b = t;  /* outTris.swap(nextTris) */
		if (n == 0) break;
	}
	if (a != cur && n > 0) memcpy(cur, a, (size_t)n * sizeof(cvert_t));
	return n;
}

int axo_clip_triangle(const float* in3x18, float* out24x18) {
	cvert_t a[MAX_CLIPPED_VERTS], b[MAX_CLIPPED_VERTS];
	memset(a, 0, sizeof a); memset(b, 0, sizeof b);
	memcpy(a, in3x18, 3 * sizeof(cvert_t));
	int n = clipTriangle(3, a, b);
	memcpy(out24x18, a, (size_t)n * sizeof(cvert_t));
	return n;
}

/* ------------------------------------------------------------------ setup: src/tiled_pipeline.cpp:53-72, 89-119 */
static inline void to_screen(v4 c, int W, int H, float* sx, float* sy, float* z) {
	float w = clampW(c.w);
	float invW = 1.0f / w;
	*z = c.z * invW;
	*sx = ((c.x * invW) + 1.0f) * 0.5f * (float)W;
	*sy = ((c.y * invW) + 1.0f) * 0.5f * (float)H;
}
static int setup_triangle(const cvert_t* v, int W, int H, tri_t* t) {
	for (int i = 0; i < 3; ++i) to_screen(v[i].clip, W, H, &t->sx[i], &t->sy[i], &t->ndcz[i]);
	float dx1 = t->sx[1] - t->sx[0], dy1 = t->sy[1] - t->sy[0];
	float dx2 = t->sx[2] - t->sx[0], dy2 = t->sy[2] - t->sy[0];
	float signedArea = dx1 * dy2 - dx2 * dy1;
	t->minx = fmin3(t->sx[0], t->sx[1], t->sx[2]); t->miny = fmin3(t->sy[0], t->sy[1], t->sy[2]);
	t->maxx = fmax3(t->sx[0], t->sx[1], t->sx[2]); t->maxy = fmax3(t->sy[0], t->sy[1], t->sy[2]);
	return signedArea < 0;  /* back face */
}
int axo_triangle_setup(const float* in3x18, int w, int h, float* out13) {
	cvert_t v[3]; tri_t t;
	memcpy(v, in3x18, sizeof v);
	int back = setup_triangle(v, w, h, &t);
	for (int i = 0; i < 3; ++i) { out13[i * 2] = t.sx[i]; out13[i * 2 + 1] = t.sy[i]; out13[6 + i] = t.ndcz[i]; }
	out13[9] = t.minx; out13[10] = t.miny; out13[11] = t.maxx; out13[12] = t.maxy;
	return back;
}

/* ------------------------------------------------------------------ textures: include/texture.hpp:12-34 */
typedef struct { const uint8_t* data; int w, h; } tex_t;
static inline v4 texel(const tex_t* t, int x, int y) {
	const uint8_t* p = t->data + ((size_t)y * (size_t)t->w + (size_t)x) * 4;
	const float inv255 = 1.0f / 255.0f;
	return V4(p[0] * inv255, p[1] * inv255, p[2] * inv255, p[3] * inv255);
}
static v4 sample_nearest(const tex_t* t, float u, float v) {
	if (!t->data) return V4(0, 0, 0, 1);
	int x = cvtt(u * (float)(t->w - 1));
	int y = cvtt(v * (float)(t->h - 1));
	x = imax(0, imin(x, t->w - 1));
	y = imax(0, imin(y, t->h - 1));
	y = t->h - 1 - y;
	return texel(t, x, y);
}
/* EXTENSION (no reference counterpart), defined in SURVEY.md §8(c) "Extension without oracle" */
static v4 sample_bilinear(const tex_t* t, float u, float v) {
	if (!t->data) return V4(0, 0, 0, 1);
	float wm = (float)(t->w - 1), hm = (float)(t->h - 1);
	float fx = u * wm, fy = v * hm;
	fx = fx > 0.0f ? fx : 0.0f; fx = fx < wm ? fx : wm;
	fy = fy > 0.0f ? fy : 0.0f; fy = fy < hm ? fy : hm;
	int x0 = (int)fx, y0 = (int)fy;
	int x1 = imin(x0 + 1, t->w - 1), y1 = imin(y0 + 1, t->h - 1);
	float tx = fx - (float)x0, ty = fy - (float)y0;
	v4 c00 = texel(t, x0, t->h - 1 - y0), c10 = texel(t, x1, t->h - 1 - y0);
	v4 c01 = texel(t, x0, t->h - 1 - y1), c11 = texel(t, x1, t->h - 1 - y1);
	v4 a = mix4(c00, c10, tx), b = mix4(c01, c11, tx);
	return mix4(a, b, ty);
}
static inline v4 sample(const tex_t* t, float u, float v, int sampler) {
	return sampler ? sample_bilinear(t, u, v) : sample_nearest(t, u, v);
}
void axo_texture_sample(const uint8_t* rgba, int w, int h, int sampler, const float* uv, int n, float* out) {
	tex_t t = {rgba, w, h};
	for (int i = 0; i < n; ++i) {
		v4 c = sample(&t, uv[i * 2], uv[i * 2 + 1], sampler);
		out[i * 4] = c.x; out[i * 4 + 1] = c.y; out[i * 4 + 2] = c.z; out[i * 4 + 3] = c.w;
	}
}

/* ------------------------------------------------------------------ shaders: include/shaders/shaders.hpp */
typedef struct {
	int kind, sampler;
	m4 model, mvp;
	m3 normal_mat;  /* mat3(transpose(inverse(model))), FlatShader :36-37 (a pure function of `model`) */
	v3 cam_pos, light_dir, light_color;
	float specular_exponent;
	tex_t tex[5];
} uniforms_t;

/* FlatShader::vertex :26-40, PhongShader::vertex :147-168, PBRShader::vertex :259-282 */
static void shader_vertex(const uniforms_t* u, const vertex_t* vin, vsout_t* o) {
	memset(o, 0, sizeof *o);
	if (u->kind == 0 || u->kind == 3) {
		o->normal = m3_mul_v3(&u->normal_mat, ld3(vin->normal));
		if (u->kind == 3) { o->uv[0] = vin->uv[0]; o->uv[1] = vin->uv[1]; }
		return;
	}
	o->uv[0] = vin->uv[0]; o->uv[1] = vin->uv[1];
	v4 wp = m4_mul_v4(&u->model, V4(vin->position[0], vin->position[1], vin->position[2], 1.0f));
	o->world = V3(wp.x, wp.y, wp.z);
	v4 t = m4_mul_v4(&u->model, V4(vin->tangent[0], vin->tangent[1], vin->tangent[2], 0.0f));
	v4 b = m4_mul_v4(&u->model, V4(vin->bitangent[0], vin->bitangent[1], vin->bitangent[2], 0.0f));
	v4 n = m4_mul_v4(&u->model, V4(vin->normal[0], vin->normal[1], vin->normal[2], 0.0f));
	o->tbn.c[0] = norm3(V3(t.x, t.y, t.z));
	o->tbn.c[1] = norm3(V3(b.x, b.y, b.z));
	o->tbn.c[2] = norm3(V3(n.x, n.y, n.z));
}

static inline v3 bary3(const float bar[3], v3 a, v3 b, v3 c) {  /* bar.x*a + bar.y*b + bar.z*c */
	return add3(add3(scl3(a, bar[0]), scl3(b, bar[1])), scl3(c, bar[2]));
}
static inline m3 bary_m3(const float bar[3], const m3* a, const m3* b, const m3* c) {
	m3 x = m3_scl(a, bar[0]), y = m3_scl(b, bar[1]), z = m3_scl(c, bar[2]);
	m3 xy = m3_add(&x, &y);
	return m3_add(&xy, &z);
}

/* FlatShader::fragment :42-57 */
static void fragment_flat(const uniforms_t* u, const float bar[3], const vsout_t* vs, float out[4]) {
	v3 n = norm3(bary3(bar, vs[0].normal, vs[1].normal, vs[2].normal));
	float intensity = clampf(dot3(neg3(u->light_dir), n), 0.0f, 1.0f);
	out[0] = out[1] = out[2] = out[3] = 1.0f * intensity;
}
/* PhongShader::fragment :170-241 */
static void fragment_phong(const uniforms_t* u, const float bar[3], const vsout_t* vs, float out[4]) {
	float uvx = bar[0] * vs[0].uv[0] + bar[1] * vs[1].uv[0] + bar[2] * vs[2].uv[0];
	float uvy = bar[0] * vs[0].uv[1] + bar[1] * vs[1].uv[1] + bar[2] * vs[2].uv[1];
	v4 nm = sample(&u->tex[1], uvx, uvy, u->sampler);
	v3 nms = norm3(sub3(scl3(V3(nm.x, nm.y, nm.z), 2.0f), V3(1.0f, 1.0f, 1.0f)));
	m3 tbn = bary_m3(bar, &vs[0].tbn, &vs[1].tbn, &vs[2].tbn);
	v3 T = tbn.c[0], N = norm3(tbn.c[2]);
	v3 Tn = norm3(sub3(T, scl3(N, dot3(N, T))));
	v3 Bn = cross3(N, Tn);
	m3 ftbn = {{Tn, Bn, N}};
	v4 albedo = sample(&u->tex[0], uvx, uvy, u->sampler);
	v3 normal = norm3(m3_mul_v3(&ftbn, nms));
	v3 fragPos = bary3(bar, vs[0].world, vs[1].world, vs[2].world);
	v3 viewDir = norm3(sub3(u->cam_pos, fragPos));
	v3 lightDir = neg3(u->light_dir);
	v3 ambient = scl3(u->light_color, 0.1f);
	float diff = maxf(dot3(normal, lightDir), 0.0f);
	v3 diffuse = scl3(u->light_color, diff);
	v3 I = neg3(lightDir);
	v3 reflectDir = sub3(I, scl3(scl3(normal, dot3(normal, I)), 2.0f));  /* glm::reflect: I - N*dot(N,I)*2 */
	float spec = powf(maxf(dot3(viewDir, reflectDir), 0.0f), u->specular_exponent * 50.0f);
	v3 specular = scl3(u->light_color, 0.5f * spec);
	v3 fc = mul3(add3(add3(ambient, diffuse), specular), V3(albedo.x, albedo.y, albedo.z));
	out[0] = fc.x; out[1] = fc.y; out[2] = fc.z; out[3] = 1.0f;
}
/* PBRShader::fragment :284-399 (only the terms that reach `color`; the re-orthogonalised basis at :316-323 is dead) */
static void fragment_pbr(const uniforms_t* u, const float bar[3], const vsout_t* vs, float out[4]) {
	const float PI = 3.14159265358979323846264338327950288f;
	float uvx = bar[0] * vs[0].uv[0] + bar[1] * vs[1].uv[0] + bar[2] * vs[2].uv[0];
	float uvy = bar[0] * vs[0].uv[1] + bar[1] * vs[1].uv[1] + bar[2] * vs[2].uv[1];
	v4 nm = sample(&u->tex[1], uvx, uvy, u->sampler);
	v3 nms = norm3(V3(nm.x * 2.0f - 1.0f, nm.y * 2.0f - 1.0f, nm.z * 2.0f - 1.0f));
	m3 tbn = bary_m3(bar, &vs[0].tbn, &vs[1].tbn, &vs[2].tbn);
	v4 al = sample(&u->tex[0], uvx, uvy, u->sampler);
	v3 albedo = V3(powf(al.x, 2.2f), powf(al.y, 2.2f), powf(al.z, 2.2f));
	float metallic = sample(&u->tex[2], uvx, uvy, u->sampler).x;
	float roughness = sample(&u->tex[3], uvx, uvy, u->sampler).x;
	float ao = sample(&u->tex[4], uvx, uvy, u->sampler).x;
	float roughness2 = roughness * roughness;
	float roughness4 = roughness2 * roughness2;
	float oneMinusMetallic = 1.0f - metallic;
	v3 normal = norm3(m3_mul_v3(&tbn, nms));
	v3 fragPos = bary3(bar, vs[0].world, vs[1].world, vs[2].world);
	v3 viewDir = norm3(sub3(u->cam_pos, fragPos));
	v3 lightDir = neg3(u->light_dir);
	v3 halfwayDir = norm3(add3(lightDir, viewDir));
	float NdotH = maxf(dot3(normal, halfwayDir), 0.0f);
	float NdotV = maxf(dot3(normal, viewDir), 0.0f);
	float NdotL = maxf(dot3(normal, lightDir), 0.0f);
	v3 c04 = V3(0.04f, 0.04f, 0.04f);
	v3 F0 = add3(c04, scl3(sub3(albedo, c04), metallic));  /* lerp: a + t*(b-a) */
	float NdotH2 = NdotH * NdotH;
	float denomPart = (NdotH2 * (roughness4 - 1.0f) + 1.0f);
	float NDF = roughness4 / (PI * denomPart * denomPart);
	float r = roughness + 1.0f;
	float k = (r * r) / 8.0f;
	float NdotV_k = NdotV * (1.0f - k) + k;
	float NdotL_k = NdotL * (1.0f - k) + k;
	float G = (NdotV / NdotV_k) * (NdotL / NdotL_k);
	float cosTheta = maxf(dot3(halfwayDir, normal), 0.0f);
	float om = 1.0f - cosTheta;
	float term = om * om; term *= term; term *= om;
	v3 one = V3(1.0f, 1.0f, 1.0f);
	v3 F = add3(F0, scl3(sub3(one, F0), term));
	v3 numerator = scl3(scl3(F, NDF), G);
	float denom = 4.0f * NdotV * NdotL + 0.0001f;
	v3 specular = V3(numerator.x / denom, numerator.y / denom, numerator.z / denom);
	v3 kD = scl3(sub3(one, F), oneMinusMetallic);
	v3 diffuse = scl3(mul3(kD, albedo), 1.0f / PI);
	v3 ambient = scl3(mul3(V3(0.03f, 0.03f, 0.03f), albedo), ao);
	v3 fc = add3(ambient, scl3(mul3(add3(diffuse, specular), u->light_color), NdotL));
	v3 fp1 = add3(fc, one);
	fc = V3(fc.x / fp1.x, fc.y / fp1.y, fc.z / fp1.z);
	const float g = 1.0f / 2.2f;
	out[0] = powf(fc.x, g); out[1] = powf(fc.y, g); out[2] = powf(fc.z, g); out[3] = 1.0f;
}
/* CutoutShader (oracle/ref_harness.cpp; not a reference shader): alpha-tested Lambert, returns 1 = discard */
static int fragment_cutout(const uniforms_t* u, const float bar[3], const vsout_t* vs, float out[4]) {
	float uvx = bar[0] * vs[0].uv[0] + bar[1] * vs[1].uv[0] + bar[2] * vs[2].uv[0];
	float uvy = bar[0] * vs[0].uv[1] + bar[1] * vs[1].uv[1] + bar[2] * vs[2].uv[1];
	v4 texl = sample(&u->tex[0], uvx, uvy, u->sampler);
	if (texl.w < 0.5f) return 1;
	v3 n = norm3(bary3(bar, vs[0].normal, vs[1].normal, vs[2].normal));
	float intensity = clampf(dot3(neg3(u->light_dir), n), 0.0f, 1.0f);
	out[0] = texl.x * intensity; out[1] = texl.y * intensity; out[2] = texl.z * intensity; out[3] = 1.0f;
	return 0;
}
/* returns the reference's `discard` flag (IShader::fragment, include/IShader.hpp:38) */
static inline int shader_fragment(const uniforms_t* u, const float bar[3], const vsout_t* vs, float out[4]) {
	if (u->kind == 0) fragment_flat(u, bar, vs, out);
	else if (u->kind == 1) fragment_phong(u, bar, vs, out);
	else if (u->kind == 2) fragment_pbr(u, bar, vs, out);
	else return fragment_cutout(u, bar, vs, out);
	return 0;
}

/* ------------------------------------------------------------------ per-tile raster: src/tiled_pipeline.cpp:427-594 */
typedef struct { float depth[TILE_SIZE * TILE_SIZE]; uint8_t color[TILE_SIZE * TILE_SIZE * 4]; } tilebuf_t;

static void raster_tri_in_tile(const uniforms_t* u, const tri_t* tri, int tsx, int tsy, int tex_, int tey, tilebuf_t* buf) {
	int startX = imax(tsx, cvtt(floorf(tri->minx)));
	int startY = imax(tsy, cvtt(floorf(tri->miny)));
	int endX = imin(tex_, cvtt(ceilf(tri->maxx)));
	int endY = imin(tey, cvtt(ceilf(tri->maxy)));
	if (startX >= endX || startY >= endY) return;
	float x0 = tri->sx[0], y0 = tri->sy[0], x1 = tri->sx[1], y1 = tri->sy[1], x2 = tri->sx[2], y2 = tri->sy[2];
	float e0_c = x1 * y2 - x2 * y1;
	float e1_c = x2 * y0 - x0 * y2;
	float e2_c = x0 * y1 - x1 * y0;
	float area = e0_c + e1_c + e2_c;
	if (area >= 0 && (double)area < 1.0E-12) return;
	float e0_a = y1 - y2, e0_b = x2 - x1;
	float e1_a = y2 - y0, e1_b = x0 - x2;
	float e2_a = y0 - y1, e2_b = x1 - x0;
	if (area < 0) {
		e0_a = -e0_a; e0_b = -e0_b; e0_c = -e0_c;
		e1_a = -e1_a; e1_b = -e1_b; e1_c = -e1_c;
		e2_a = -e2_a; e2_b = -e2_b; e2_c = -e2_c;
		area = -area;
	}
	float invArea = 1.0f / area;
	for (int py = startY; py < endY; ++py) {
		float py_center = py + 0.5f;
		float row_e0 = e0_a * (startX + 0.5f) + e0_b * py_center + e0_c;
		float row_e1 = e1_a * (startX + 0.5f) + e1_b * py_center + e1_c;
		float row_e2 = e2_a * (startX + 0.5f) + e2_b * py_center + e2_c;
		for (int px = startX; px < endX; px += 8) {
			for (int i = 0; i < 8; ++i) {  /* lanes of the __m256; pxOffsets = 0..7 */
				float c0 = row_e0 + e0_a * (float)i;
				float c1 = row_e1 + e1_a * (float)i;
				float c2 = row_e2 + e2_a * (float)i;
				if (!(c0 >= 0.f && c1 >= 0.f && c2 >= 0.f)) continue;  /* _CMP_GE_OQ */
				int curX = px + i;
				if (curX >= endX) break;
				float bar[3] = {c0 * invArea, c1 * invArea, c2 * invArea};
				float z = tri->ndcz[0] * bar[0] + tri->ndcz[1] * bar[1] + tri->ndcz[2] * bar[2];
				int idx = (py - tsy) * TILE_SIZE + (curX - tsx);
				if (z < buf->depth[idx]) {
					float col[4];
					if (shader_fragment(u, bar, tri->vs, col)) continue;  /* :571-577: discard keeps depth and colour */
					buf->depth[idx] = z;
					for (int c = 0; c < 4; ++c) buf->color[idx * 4 + c] = (uint8_t)cvtt(clampf(col[c], 0.0f, 1.0f) * 255.0f);
				}
			}
			row_e0 += e0_a * 8.0f;
			row_e1 += e1_a * 8.0f;
			row_e2 += e2_a * 8.0f;
		}
	}
}

/* ------------------------------------------------------------------ drawMesh: src/tiled_pipeline.cpp:143-322 */
typedef struct {
	const uniforms_t* u;
	const tri_t* tris;
	const uint32_t* bin_start;  /* per tile, into bin_items */
	const uint32_t* bin_items;
	int W, H, ntx, nty;
	uint8_t* color; float* depth;
	int next_tile;              /* atomic work counter (the reference's batch dispenser, :285-301) */
} job_t;

static void process_tile(const job_t* j, int tile, tilebuf_t* buf) {
	uint32_t b0 = j->bin_start[tile], b1 = j->bin_start[tile + 1];
	if (b0 == b1) return;
	int tx = tile % j->ntx, ty = tile / j->ntx;
	int tsx = tx * TILE_SIZE, tsy = ty * TILE_SIZE;
	int tex_ = imin((tx + 1) * TILE_SIZE, j->W), tey = imin((ty + 1) * TILE_SIZE, j->H);
	for (int i = 0; i < TILE_SIZE * TILE_SIZE; ++i) buf->depth[i] = INFINITY;  /* InlinedBuffers::initialize, include/tiled_pipeline.hpp:48-51 */
	memset(buf->color, 0, sizeof buf->color);
	for (uint32_t k = b0; k < b1; ++k) raster_tri_in_tile(j->u, &j->tris[j->bin_items[k]], tsx, tsy, tex_, tey, buf);
	/* mergeTileResults :1125-1181: strict tileZ < fbZ (ordered), R<->B swizzle into B,G,R,A */
	for (int y = 0; y < tey - tsy; ++y)
		for (int x = 0; x < tex_ - tsx; ++x) {
			int li = y * TILE_SIZE + x;
			size_t gi = (size_t)(tsy + y) * (size_t)j->W + (size_t)(tsx + x);
			if (buf->depth[li] < j->depth[gi]) {
				j->depth[gi] = buf->depth[li];
				j->color[gi * 4 + 0] = buf->color[li * 4 + 2];
				j->color[gi * 4 + 1] = buf->color[li * 4 + 1];
				j->color[gi * 4 + 2] = buf->color[li * 4 + 0];
				j->color[gi * 4 + 3] = buf->color[li * 4 + 3];
			}
		}
}
static void* tile_worker(void* p) {
	job_t* j = (job_t*)p;
	tilebuf_t buf;
	for (;;) {
		int t = __atomic_fetch_add(&j->next_tile, 1, __ATOMIC_RELAXED);
		if (t >= j->ntx * j->nty) break;
		process_tile(j, t, &buf);
	}
	return NULL;
}

/* One reference drawMesh over faces [f0, f1). */
static int draw_chunk(const uniforms_t* u, const vertex_t* verts, const uint32_t* idx, uint64_t f0, uint64_t f1,
                      int W, int H, int threads, uint8_t* color, float* depth) {
	size_t cap = (size_t)(f1 - f0) + 16, nt = 0;
	tri_t* tris = (tri_t*)malloc(cap * sizeof(tri_t));
	if (!tris) return -1;
	cvert_t a[MAX_CLIPPED_VERTS], b[MAX_CLIPPED_VERTS];
	/* transform + clip + cull + setup (:195-240) */
	for (uint64_t f = f0; f < f1; ++f) {
		memset(a, 0, sizeof a); memset(b, 0, sizeof b);
		for (int k = 0; k < 3; ++k) {
			const vertex_t* v = &verts[idx[f * 3 + k]];
			a[k].v = *v;
			a[k].clip = m4_mul_v4(&u->mvp, V4(v->position[0], v->position[1], v->position[2], 1.0f));
		}
		int n = clipTriangle(3, a, b);
		for (int j = 0; j + 2 < n; j += 3) {
			if (nt == cap) {
				cap *= 2;
				tri_t* nt_ = (tri_t*)realloc(tris, cap * sizeof(tri_t));
				if (!nt_) { free(tris); return -1; }
				tris = nt_;
			}
			tri_t* t = &tris[nt];
			if (setup_triangle(&a[j], W, H, t)) continue;  /* isBackface :225 */
			for (int k = 0; k < 3; ++k) shader_vertex(u, &a[j + k].v, &t->vs[k]);
			++nt;
		}
	}
	/* binTrianglesToTiles (:368-393) as count -> prefix -> fill, preserving triangle order per tile */
	int ntx = (W + TILE_SIZE - 1) / TILE_SIZE, nty = (H + TILE_SIZE - 1) / TILE_SIZE;
	size_t ntiles = (size_t)ntx * nty;
	uint32_t* start = (uint32_t*)calloc(ntiles + 1, sizeof(uint32_t));   /* exclusive prefix of per-tile counts */
	uint32_t* cursor = (uint32_t*)calloc(ntiles + 1, sizeof(uint32_t));
	uint32_t* items = NULL;
	for (int pass = 0; pass < 2; ++pass) {
		for (size_t i = 0; i < nt; ++i) {
			const tri_t* t = &tris[i];
			int stx = imax(0, cvtt(t->minx) / TILE_SIZE), sty = imax(0, cvtt(t->miny) / TILE_SIZE);
			int etx = imin(ntx - 1, cvtt(t->maxx) / TILE_SIZE), ety = imin(nty - 1, cvtt(t->maxy) / TILE_SIZE);
			for (int ty = sty; ty <= ety; ++ty)
				for (int tx = stx; tx <= etx; ++tx) {
					int tsx = tx * TILE_SIZE, tsy = ty * TILE_SIZE;
					int tex_ = imin((tx + 1) * TILE_SIZE, W), tey = imin((ty + 1) * TILE_SIZE, H);
					/* triangleIntersectsTile :356-366 */
					if (t->maxx < (float)tsx || t->minx > (float)tex_ || t->maxy < (float)tsy || t->miny > (float)tey) continue;
					size_t ti = (size_t)ty * ntx + tx;
					if (pass == 0) cursor[ti]++;
					else items[cursor[ti]++] = (uint32_t)i;
				}
		}
		if (pass == 0) {
			for (size_t i = 0; i < ntiles; ++i) { start[i + 1] = start[i] + cursor[i]; cursor[i] = start[i]; }
			items = (uint32_t*)malloc(((size_t)start[ntiles] + 1) * sizeof(uint32_t));
		}
	}
	free(cursor);
	job_t job = {u, tris, start, items, W, H, ntx, nty, color, depth, 0};
	if (threads <= 1) tile_worker(&job);
	else {
		if (threads > 256) threads = 256;
		pthread_t th[256];
		for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, tile_worker, &job);
		for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
	}
	free(items); free(start); free(tris);
	return 0;
}

int axo_render(const axo_scene* sc, const float* vertices, uint64_t n_verts, const uint32_t* indices,
               uint64_t n_faces, uint8_t* color_inout, float* depth_inout, double* seconds_out) {
	(void)n_verts;
	uniforms_t u;
	memset(&u, 0, sizeof u);
	u.kind = sc->shader_kind;
	u.sampler = sc->sampler;
	u.model = ld_m4(sc->model);
	m4 vp = ld_m4(sc->view_proj);
	u.mvp = m4_mul(&vp, &u.model);  /* :149 */
	m4 inv = m4_inverse(&u.model);
	/* mat3(transpose(inverse(model))): column c of the transpose = row c of the inverse */
	u.normal_mat.c[0] = V3(inv.c[0].x, inv.c[1].x, inv.c[2].x);
	u.normal_mat.c[1] = V3(inv.c[0].y, inv.c[1].y, inv.c[2].y);
	u.normal_mat.c[2] = V3(inv.c[0].z, inv.c[1].z, inv.c[2].z);
	u.cam_pos = ld3(sc->cam_pos);
	u.light_dir = ld3(sc->light_dir);
	u.light_color = ld3(sc->light_color);
	u.specular_exponent = sc->specular_exponent;
	for (int i = 0; i < 5; ++i) { u.tex[i].data = sc->tex[i]; u.tex[i].w = sc->tex_w[i]; u.tex[i].h = sc->tex_h[i]; }
	if (u.kind < 0 || u.kind > 3) return -2;
	if ((u.kind == 1 || u.kind == 2) && (!u.tex[0].data || !u.tex[1].data)) return -2;
	if (u.kind == 3 && !u.tex[0].data) return -2;
	if (u.kind == 2 && (!u.tex[2].data || !u.tex[3].data || !u.tex[4].data)) return -2;
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	/* Chunked draws composite through the strict depth test exactly like one draw; chunking bounds memory. */
	const uint64_t chunk = 1u << 16;
	for (uint64_t f0 = 0; f0 < n_faces; f0 += chunk) {
		uint64_t f1 = f0 + chunk < n_faces ? f0 + chunk : n_faces;
		int rc = draw_chunk(&u, (const vertex_t*)vertices, indices, f0, f1, sc->width, sc->height, sc->threads, color_inout, depth_inout);
		if (rc) return rc;
	}
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if (seconds_out) *seconds_out = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
	return 0;
}
