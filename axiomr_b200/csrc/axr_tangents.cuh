// Tangent / bitangent generation on the device — Mesh::calculateTangentBitangent (reference src/mesh.cpp:222-298), the step of
// mesh ingestion that is quadratic-ish on the CPU (two std::map<uint32_t, vec3> updated per face corner). SURVEY.md §8(f) rank 3.
//
// The reference accumulates per-vertex sums in face order, corner order (float addition is not associative), so the device
// version builds a vertex -> corner adjacency (count, exclusive scan, unordered fill, per-vertex sort of the corner ids) and
// every vertex then adds its faces' contributions in exactly that order. All arithmetic is binary32 without contraction
// (the library is compiled with -fmad=false), in the reference's operation order, quirks included:
//   * |denominator| < 1e-8 -> fallback tangent (0,1,0) if |n.x| > 0.8 else (1,0,0), fallback bitangent cross(n, fallback)
//   * bitangent flipped when dot(cross(n0, tangent), bitangent) < 0, with n0 the FIRST corner's normal
//   * vertices no face touches: tangent (1,0,0); the handedness test reads a zero bitangent sum
//   * the `length() < 1e-8` fallbacks never fire (glm's vec3::length() is the component count)
#pragma once
#include <cub/device/device_scan.cuh>

#include "axr_math.cuh"

namespace axr {

struct FaceTB { float t[3]; float b[3]; };

__device__ __forceinline__ v3 ld_v3(const float* p) { return V3(p[0], p[1], p[2]); }

// v8: V x 8 floats (position3, uv2, normal3)
__global__ void k_tan_faces(const float* __restrict__ v8, const unsigned* __restrict__ idx, unsigned long long n_faces,
                            FaceTB* __restrict__ out, unsigned* __restrict__ degree) {
	unsigned long long f = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= n_faces) return;
	const unsigned a = idx[f * 3], b = idx[f * 3 + 1], c = idx[f * 3 + 2];
	const float *pa = v8 + (size_t)a * 8, *pb = v8 + (size_t)b * 8, *pc = v8 + (size_t)c * 8;
	const v3 e1 = ld_v3(pb) - ld_v3(pa), e2 = ld_v3(pc) - ld_v3(pa);
	const float d1x = pb[3] - pa[3], d1y = pb[4] - pa[4], d2x = pc[3] - pa[3], d2y = pc[4] - pa[4];
	const float den = d1x * d2y - d2x * d1y;
	const v3 n0 = ld_v3(pa + 5);
	v3 tan, bit;
	if (fabsf(den) < 1e-8f) {
		tan = (fabsf(n0.x) > 0.8f) ? V3(0.f, 1.f, 0.f) : V3(1.f, 0.f, 0.f);
		bit = cross(n0, tan);
	} else {
		const float fi = 1.0f / den;
		tan = (e1 * d2y - e2 * d1y) * fi;
		bit = (e1 * (-d2x) + e2 * d1x) * fi;
		if (dot(cross(n0, tan), bit) < 0.0f) bit = -bit;
	}
	FaceTB r;
	r.t[0] = tan.x; r.t[1] = tan.y; r.t[2] = tan.z;
	r.b[0] = bit.x; r.b[1] = bit.y; r.b[2] = bit.z;
	out[f] = r;
	atomicAdd(degree + a, 1u);
	atomicAdd(degree + b, 1u);
	atomicAdd(degree + c, 1u);
}

__global__ void k_tan_fill(const unsigned* __restrict__ idx, unsigned long long n_faces, const unsigned* __restrict__ start,
                           unsigned* __restrict__ cursor, unsigned* __restrict__ corners) {
	unsigned long long f = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= n_faces) return;
	for (int k = 0; k < 3; ++k) {
		const unsigned v = idx[f * 3 + k];
		corners[start[v] + atomicAdd(cursor + v, 1u)] = (unsigned)(f * 3 + k);
	}
}

__global__ void k_tan_vertices(const float* __restrict__ v8, unsigned long long n_verts, const unsigned* __restrict__ start,
                               unsigned* __restrict__ corners, const FaceTB* __restrict__ ftb, float* __restrict__ out14) {
	unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= n_verts) return;
	const unsigned s0 = start[v], s1 = start[v + 1];
	// corner ids into face order, corner order. Insertion sort for the valences real meshes have; a pole or fan vertex with thousands
	// of incident faces would make that one thread quadratic, so long lists take an in-place heapsort (same result, O(n log n)).
	const unsigned cnt = s1 - s0;
	unsigned* c = corners + s0;
	if (cnt <= 32u) {
		for (unsigned i = 1; i < cnt; ++i) {
			const unsigned key = c[i];
			unsigned j = i;
			while (j > 0 && c[j - 1] > key) { c[j] = c[j - 1]; --j; }
			c[j] = key;
		}
	} else {
		auto sift = [&](unsigned root, unsigned end) {  // max-heap on c[0, end)
			const unsigned key = c[root];
			for (;;) {
				unsigned child = 2u * root + 1u;
				if (child >= end) break;
				if (child + 1u < end && c[child + 1u] > c[child]) ++child;
				if (!(c[child] > key)) break;
				c[root] = c[child];
				root = child;
			}
			c[root] = key;
		};
		for (unsigned i = cnt / 2u; i-- > 0;) sift(i, cnt);
		for (unsigned end = cnt - 1u; end > 0; --end) {
			const unsigned top = c[0];
			c[0] = c[end]; c[end] = top;
			sift(0, end);
		}
	}
	v3 ts = V3(0.f, 0.f, 0.f), bs = V3(0.f, 0.f, 0.f);
	for (unsigned i = s0; i < s1; ++i) {
		const FaceTB r = ftb[corners[i] / 3u];
		ts = ts + V3(r.t[0], r.t[1], r.t[2]);
		bs = bs + V3(r.b[0], r.b[1], r.b[2]);
	}
	const float* p = v8 + (size_t)v * 8;
	const v3 n = ld_v3(p + 5);
	v3 t = (s1 > s0) ? ts : V3(1.f, 0.f, 0.f);
	t = t - n * dot(n, t);
	t = normalize(t);
	v3 b = cross(n, t);
	const float hand = (dot(bs, b) < 0.0f) ? -1.0f : 1.0f;
	b = b * hand;
	float* o = out14 + (size_t)v * 14;
	for (int i = 0; i < 8; ++i) o[i] = p[i];
	o[8] = t.x; o[9] = t.y; o[10] = t.z;
	o[11] = b.x; o[12] = b.y; o[13] = b.z;
}

}  // namespace axr
