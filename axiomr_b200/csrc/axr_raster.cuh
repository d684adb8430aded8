// Clip, triangle setup, coverage and visibility keys — device restatement of the reference's
// per-triangle / per-pixel arithmetic, arranged for one-thread-per-item GPU execution.
// Citations are relative to the reference tree.
#pragma once
#include "axr_math.cuh"

namespace axr {

constexpr int REF_TILE = 16;          // reference include/tiled_pipeline.hpp:28 (coverage arithmetic depends on it)
constexpr int MAX_CLIPPED_VERTS = 24; // reference include/pipeline.hpp:17
constexpr unsigned long long KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;

// AR::Vertex without the position (reference include/mesh.hpp:9-18): uv2, normal3, tangent3, bitangent3 (+1 pad) = 48 B
struct VAttr { float uv[2]; float n[3]; float t[3]; float b[3]; float pad; };
static_assert(sizeof(VAttr) == 48, "VAttr layout");

// ------------------------------------------------------------------ clampW + perspective divide
// reference src/tiled_pipeline.cpp:15-23
__device__ __forceinline__ float clampW(float w) {
	const float tiny = 1e-6f;
	if (fabsf(w) < tiny) w = (w < 0.0f) ? -tiny : tiny;
	return w;
}
// reference src/tiled_pipeline.cpp:57-66 (Triangle ctor) == :96-106 (isBackface): same expression order
__device__ __forceinline__ void to_screen(v4 c, float fW, float fH, float& sx, float& sy, float& z) {
	float invW = 1.0f / clampW(c.w);
	z = c.z * invW;
	sx = ((c.x * invW) + 1.0f) * 0.5f * fW;
	sy = ((c.y * invW) + 1.0f) * 0.5f * fH;
}
// reference src/pipeline.cpp:274-285; bit k set <=> vertex is NOT inside plane k (inside <=> d >= 0, NaN counts as not inside)
__device__ __forceinline__ unsigned clip_code(v4 c) {
	unsigned code = 0;
	code |= ((c.x + c.w) >= 0.f) ? 0u : 1u;
	code |= ((c.w - c.x) >= 0.f) ? 0u : 2u;
	code |= ((c.y + c.w) >= 0.f) ? 0u : 4u;
	code |= ((c.w - c.y) >= 0.f) ? 0u : 8u;
	code |= ((c.z + c.w) >= 0.f) ? 0u : 16u;
	code |= ((c.w - c.z) >= 0.f) ? 0u : 32u;
	return code;
}
// "Safely outside" code, bit k set <=> d_k < -(1e-4*(|w|+|c|) + 1e-6). If all three vertices of a face share such a bit,
// clipTriangle (reference src/pipeline.cpp:176-228) provably returns nothing: every vertex the earlier planes create is a
// convex combination x*(1-t) + y*t, t in [0,1], of vertices that are outside plane k by that margin; the rounding of the
// products, sums and of 1-t moves d_k by a few 2^-24 relative to |w|+|c| per generation (at most five generations), four
// orders of magnitude below the margin, so every sub-triangle reaches plane k with d < 0 on all vertices and is dropped
// (:311-313). Without the margin the claim is false: d can round to exactly 0 or fall inside the |denom| < 1e-7 -> t = 0.5
// fallback (:343-344), so faces that are outside by less than the margin take the exact slow path instead.
__device__ __forceinline__ unsigned clip_code_safe_out(v4 c) {
	const float aw = fabsf(c.w);
	const float mx = 1e-4f * (aw + fabsf(c.x)) + 1e-6f, my = 1e-4f * (aw + fabsf(c.y)) + 1e-6f, mz = 1e-4f * (aw + fabsf(c.z)) + 1e-6f;
	unsigned code = 0;
	code |= ((c.x + c.w) < -mx) ? 1u : 0u;
	code |= ((c.w - c.x) < -mx) ? 2u : 0u;
	code |= ((c.y + c.w) < -my) ? 4u : 0u;
	code |= ((c.w - c.y) < -my) ? 8u : 0u;
	code |= ((c.z + c.w) < -mz) ? 16u : 0u;
	code |= ((c.w - c.z) < -mz) ? 32u : 0u;
	return code;
}
__device__ __forceinline__ float dist_func(v4 v, int plane) {
	switch (plane) {
	case 0: return v.x + v.w;
	case 1: return v.w - v.x;
	case 2: return v.y + v.w;
	case 3: return v.w - v.y;
	case 4: return v.z + v.w;
	default: return v.w - v.z;
	}
}

// ------------------------------------------------------------------ McGuire clip (reference src/pipeline.cpp:176-370)
// Vertex payloads: position-only for the visibility pass, full attributes for the shading pass.
struct ClipPos { v4 clip; };
struct ClipFull { v4 clip; v3 pos; float uv[2]; v3 n, t, b; };

__device__ __forceinline__ ClipPos interpolate(const ClipPos& a, const ClipPos& b, float t) {
	ClipPos o;
	o.clip = mix(a.clip, b.clip, t);
	return o;
}
// reference src/pipeline.cpp:243-272
__device__ __forceinline__ ClipFull interpolate(const ClipFull& a, const ClipFull& b, float t) {
	ClipFull o;
	o.pos = mix(a.pos, b.pos, t);
	o.n = normalize(mix(a.n, b.n, t));
	o.t = normalize(mix(a.t, b.t, t));
	o.b = normalize(mix(a.b, b.b, t));
	o.clip = mix(a.clip, b.clip, t);
	float w0 = a.clip.w, w1 = b.clip.w;
	float u0x = a.uv[0] * w0, u0y = a.uv[1] * w0, u1x = b.uv[0] * w1, u1y = b.uv[1] * w1;
	float ux = u0x * (1.0f - t) + u1x * t, uy = u0y * (1.0f - t) + u1y * t;
	float iw = mixf(w0, w1, t);
	o.uv[0] = ux / iw;
	o.uv[1] = uy / iw;
	return o;
}
template <typename CV>
__device__ __forceinline__ void swapcv(CV& a, CV& b) { CV t = a; a = b; b = t; }

// reference src/pipeline.cpp:302-370
template <typename CV>
__device__ int clip_single_plane(int plane, CV& v0, CV& v1, CV& v2, CV& v3) {
	float d0 = dist_func(v0.clip, plane), d1 = dist_func(v1.clip, plane), d2 = dist_func(v2.clip, plane);
	if (d0 < 0.f && d1 < 0.f && d2 < 0.f) return 0;
	if (d0 >= 0.f && d1 >= 0.f && d2 >= 0.f) { v3 = v0; return 3; }
	float td;
	if (d1 >= 0.f && !(d0 >= 0.f)) {
		swapcv(v0, v1); td = d0; d0 = d1; d1 = td;
		swapcv(v1, v2); td = d1; d1 = d2; d2 = td;
	} else if (d2 >= 0.f && !(d1 >= 0.f)) {
		swapcv(v2, v1); td = d2; d2 = d1; d1 = td;
		swapcv(v1, v0); td = d1; d1 = d0; d0 = td;
	}
	float denom02 = d0 - d2;
	float t02 = (fabsf(denom02) < 1e-7f) ? 0.5f : (d0 / denom02);
	v3 = interpolate(v0, v2, t02);
	if (d1 >= 0.f) {
		float denom12 = d1 - d2;
		float t12 = (fabsf(denom12) < 1e-7f) ? 0.5f : (d1 / denom12);
		v2 = interpolate(v1, v2, t12);
		return 4;
	}
	float denom01 = d0 - d1;
	float t01 = (fabsf(denom01) < 1e-7f) ? 0.5f : (d0 / denom01);
	v1 = interpolate(v0, v1, t01);
	v2 = v3;
	return 3;
}
// Planes the clipper has to visit for a triangle with clip-space vertices c0, c1, c2: bit k clear <=> all three vertices are inside
// plane k by more than 1e-4 * (largest |w| + |c| among them) + 1e-6. Every vertex the other planes create is a chain of at most
// six convex combinations x*(1-t) + y*t, t in [0,1], of such vertices; the rounding of the products, the sums and of 1-t moves d_k
// by a few 2^-24 of that largest magnitude per generation, three orders of magnitude below the margin, so at plane k every
// sub-triangle has d >= 0 on all vertices and clip_single_plane hands it on unchanged (:322-325): skipping the plane is exact.
__device__ __forceinline__ unsigned clip_planes_needed(v4 c0, v4 c1, v4 c2) {
	const float aw = fmaxf(fmaxf(fabsf(c0.w), fabsf(c1.w)), fabsf(c2.w));
	const float mx = 1e-4f * (aw + fmaxf(fmaxf(fabsf(c0.x), fabsf(c1.x)), fabsf(c2.x))) + 1e-6f;
	const float my = 1e-4f * (aw + fmaxf(fmaxf(fabsf(c0.y), fabsf(c1.y)), fabsf(c2.y))) + 1e-6f;
	const float mz = 1e-4f * (aw + fmaxf(fmaxf(fabsf(c0.z), fabsf(c1.z)), fabsf(c2.z))) + 1e-6f;
	unsigned need = 0;
	need |= ((c0.x + c0.w) > mx && (c1.x + c1.w) > mx && (c2.x + c2.w) > mx) ? 0u : 1u;
	need |= ((c0.w - c0.x) > mx && (c1.w - c1.x) > mx && (c2.w - c2.x) > mx) ? 0u : 2u;
	need |= ((c0.y + c0.w) > my && (c1.y + c1.w) > my && (c2.y + c2.w) > my) ? 0u : 4u;
	need |= ((c0.w - c0.y) > my && (c1.w - c1.y) > my && (c2.w - c2.y) > my) ? 0u : 8u;
	need |= ((c0.z + c0.w) > mz && (c1.z + c1.w) > mz && (c2.z + c2.w) > mz) ? 0u : 16u;
	need |= ((c0.w - c0.z) > mz && (c1.w - c1.z) > mz && (c2.w - c2.z) > mz) ? 0u : 32u;
	return need;  // NaN anywhere fails the compares: the plane is visited
}

// reference src/pipeline.cpp:176-228. bufA holds 3 vertices on entry. Returns the vertex count and the
// buffer (bufA or bufB) holding the result in *out. planes: clip_planes_needed() of the three vertices (0x3f = all six).
template <typename CV>
__device__ __noinline__ int clip_triangle(CV* bufA, CV* bufB, CV** out, unsigned planes) {
	CV* a = bufA;
	CV* b = bufB;
	int n = 3;
	for (int plane = 0; plane < 6; ++plane) {
		if (!((planes >> plane) & 1u)) continue;
		int cnt = 0;
		for (int i = 0; i < n; i += 3) {
			CV q;
			int r = clip_single_plane(plane, a[i], a[i + 1], a[i + 2], q);
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
		CV* t = a; a = b; b = t;
		if (n == 0) break;
	}
	*out = a;
	return n;
}

// ------------------------------------------------------------------ triangle setup (reference src/tiled_pipeline.cpp:436-486)
struct Setup {
	float a0, b0, c0, a1, b1, c1, a2, b2, c2;  // edge functions after the area-sign flip
	float inv_area;
	float z0, z1, z2;
	int fminx;              // (int)floor(minX), unclamped: the row start inside each reference tile depends on it
	int X0, X1, Y0, Y1;     // pixel box [X0,X1) x [Y0,Y1) clamped to the frame / band
};

// Back-face test, reference src/tiled_pipeline.cpp:107-118: cull <=> dx1*dy2 - dx2*dy1 < 0
__device__ __forceinline__ bool is_backface(float x0, float y0, float x1, float y1, float x2, float y2) {
	float dx1 = x1 - x0, dy1 = y1 - y0, dx2 = x2 - x0, dy2 = y2 - y0;
	float signedArea = dx1 * dy2 - dx2 * dy1;
	return signedArea < 0;
}

// Edge equations, area sign flip and 1/area (reference src/tiled_pipeline.cpp:450-486). Returns false on the degenerate-area return.
__device__ __forceinline__ bool setup_edges(float x0, float y0, float x1, float y1, float x2, float y2, float z0, float z1, float z2, Setup& s) {
	float e0_c = x1 * y2 - x2 * y1;
	float e1_c = x2 * y0 - x0 * y2;
	float e2_c = x0 * y1 - x1 * y0;
	float area = e0_c + e1_c + e2_c;
	if (area >= 0 && (double)area < 1.0E-12) return false;  // :460 (float promoted to double for the compare)
	float e0_a = y1 - y2, e0_b = x2 - x1;
	float e1_a = y2 - y0, e1_b = x0 - x2;
	float e2_a = y0 - y1, e2_b = x1 - x0;
	if (area < 0) {
		e0_a = -e0_a; e0_b = -e0_b; e0_c = -e0_c;
		e1_a = -e1_a; e1_b = -e1_b; e1_c = -e1_c;
		e2_a = -e2_a; e2_b = -e2_b; e2_c = -e2_c;
		area = -area;
	}
	s.a0 = e0_a; s.b0 = e0_b; s.c0 = e0_c;
	s.a1 = e1_a; s.b1 = e1_b; s.c1 = e1_c;
	s.a2 = e2_a; s.b2 = e2_b; s.c2 = e2_c;
	s.inv_area = 1.0f / area;
	s.z0 = z0; s.z1 = z1; s.z2 = z2;
	return true;
}

// Returns false when the reference would draw nothing for this triangle (empty box or the degenerate-area return).
// IN_FRAME: the three vertices passed all six clip planes (clip code 0), so their screen coordinates are finite and within a rounding
// error of [0, W] x [0, H]: the float -> int conversions cannot leave the int range and need no cvttss2si emulation.
template <bool IN_FRAME = false>
__device__ __forceinline__ bool setup_triangle(float x0, float y0, float x1, float y1, float x2, float y2, float z0,
                                               float z1, float z2, int W, int y_lo, int y_hi, Setup& s) {
	float minx = min3f(x0, x1, x2), miny = min3f(y0, y1, y2);
	float maxx = max3f(x0, x1, x2), maxy = max3f(y0, y1, y2);
	// Union over the reference's 16x16 tiles of [max(tile.startX, floor(minX)), min(tile.endX, ceil(maxX))) (:436-440)
	auto to_int = [](float v) { return IN_FRAME ? (int)v : cvtt(v); };
	s.fminx = to_int(floorf(minx));
	s.X0 = max(0, s.fminx);
	s.X1 = min(W, to_int(ceilf(maxx)));
	s.Y0 = max(y_lo, to_int(floorf(miny)));
	s.Y1 = min(y_hi, to_int(ceilf(maxy)));
	if (s.X0 >= s.X1 || s.Y0 >= s.Y1) return false;
	return setup_edges(x0, y0, x1, y1, x2, y2, z0, z1, z2, s);
}

// Coverage of pixel (px,py), px in [X0,X1), as a closed form of the reference's AVX2 loop (:503-537):
// inside the pixel's 16x16 reference tile the row starts at startX = max(tile.startX, floor(minX)), the edge value
// is advanced by a*8.0f for the second group of eight, then by a*(float)i for lane i. No FMA anywhere.
__device__ __forceinline__ bool coverage(const Setup& s, int px, int py, float& c0, float& c1, float& c2) {
	int startX = max(px & ~(REF_TILE - 1), s.fminx);
	float sxc = (float)startX + 0.5f;
	float pyc = (float)py + 0.5f;
	float r0 = s.a0 * sxc + s.b0 * pyc + s.c0;
	float r1 = s.a1 * sxc + s.b1 * pyc + s.c1;
	float r2 = s.a2 * sxc + s.b2 * pyc + s.c2;
	int d = px - startX;
	if (d >= 8) {
		r0 = r0 + s.a0 * 8.0f;
		r1 = r1 + s.a1 * 8.0f;
		r2 = r2 + s.a2 * 8.0f;
		d -= 8;
	}
	float fi = (float)d;
	c0 = r0 + s.a0 * fi;
	c1 = r1 + s.a1 * fi;
	c2 = r2 + s.a2 * fi;
	return c0 >= 0.f && c1 >= 0.f && c2 >= 0.f;  // _CMP_GE_OQ: NaN fails
}
// :555-562
__device__ __forceinline__ float interp_z(const Setup& s, float c0, float c1, float c2, float& al, float& be, float& ga) {
	al = c0 * s.inv_area;
	be = c1 * s.inv_area;
	ga = c2 * s.inv_area;
	return s.z0 * al + s.z1 * be + s.z2 * ga;
}

// ------------------------------------------------------------------ visibility key
// The reference keeps, per pixel, the first triangle in m_Triangles order with the strictly smallest z
// (:569 `z < depth`, depth starts at +inf, triangles visited in order :403). That is argmin over (z, ordinal), so a
// 64-bit atomicMin on (orderable(z) << 32 | ordinal) reproduces it for any processing order. -0 and +0 compare equal
// in the reference, so the key canonicalises -0; z = +inf or NaN never passes `z < +inf`.
__device__ __forceinline__ bool z_draws(float z) { return z < INFINITY; }
__device__ __forceinline__ unsigned long long make_key(float z, unsigned ordinal) {
	unsigned b = __float_as_uint(z + 0.0f);
	// negative: ~b, else b | 0x80000000 — as one shift and one xor: the mask is all ones for a set sign bit, the sign bit alone otherwise
	b ^= (unsigned)((int)b >> 31) | 0x80000000u;
	return ((unsigned long long)b << 32) | ordinal;
}

}  // namespace axr
