// TEST INFRASTRUCTURE ONLY — a SIMT interpreter that lets the CPU test suite execute the *unmodified* CUDA sources of
// axiomr_b200/csrc (kernels and the C ABI) without a GPU, so that kernel logic (indexing, warp collectives, barriers, bins,
// depth peeling, error paths) is exercised by `pytest -m "not gpu"` every round and kernel changes can be checked against the
// oracle before any GPU time is spent. It is NOT a product path and NOT a fallback: nothing under axiomr_b200/, bench.py or
// __graft_entry__.py builds, loads or mentions it; the library it produces (tests/simt/_build/libaxr_simt.so) is only ever
// loaded by tests/test_simt_kernels.py in a subprocess. Execution is serial and ~1000x slower than one CPU core running the
// oracle; it proves nothing about performance.
//
// This header shadows <cuda_runtime.h> (tests/simt/include comes first on the include path of tests/simt/build.py):
//   * execution-space / launch qualifiers become no-ops, __shared__ becomes `static` (CTAs run one after the other),
//   * threadIdx is the running fiber's index; blockIdx / blockDim / gridDim are set per CTA,
//   * every thread of a CTA is a fiber (own stack, hand-written context switch); __syncthreads, __syncwarp and the *_sync warp
//     collectives are rendezvous points among the fibers of the CTA / warp, with a watchdog that aborts on a deadlock and a check
//     that all lanes of a warp meet at the same kind of collective,
//   * the CUDA runtime calls the C ABI layer makes are mapped onto malloc / memcpy; device allocations carry canary zones that
//     are checked at every synchronisation and free (out-of-bounds kernel stores abort the test) and are filled with 0xCD so
//     that reliance on zero-initialised device memory shows,
//   * `kernel<<<grid, block, smem, stream>>>(args)` is rewritten textually by build.py into simt::launch_cfg(...).run(kernel, args).
#pragma once
// Standard headers first: the qualifier macros below (notably __noinline__) must not leak into them.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#define __host__
#define __device__
#define __global__
#define __shared__ static
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))

// ------------------------------------------------------------------------------------------------ vector types
struct uint3 { unsigned x, y, z; };
struct dim3 {
	unsigned x, y, z;
	dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(4) uchar4 { unsigned char x, y, z, w; };
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { uchar4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

// ------------------------------------------------------------------------------------------------ interpreter core (simt.cpp)
namespace simt {
struct Fiber {
	void* sp;
	uint3 tid;
	unsigned linear, lane, warp;
	bool done;
};
extern Fiber* g_cur;
enum Op { OP_SYNCWARP = 1, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_BALLOT, OP_MATCH, OP_REDUCE };
// All lanes of `mask` that are still running deposit `v`; returns once every one of them has, with the 32 deposited values in
// out[] and the set of participating lanes as the return value.
unsigned warp_exchange(unsigned mask, int op, unsigned long long v, unsigned long long out[32]);
void cta_barrier();
void run_grid(dim3 grid, dim3 block, const std::function<void()>& thread_body, const char* name);

struct launch_cfg {
	dim3 g, b;
	launch_cfg(dim3 g_, dim3 b_, size_t = 0, void* = nullptr) : g(g_), b(b_) {}
	template <typename... P, typename... A>
	void run(void (*kern)(P...), A&&... a) const {
		std::tuple<std::decay_t<P>...> params(std::forward<A>(a)...);  // by value, converted to the kernel's parameter types
		run_grid(g, b, [&] { std::apply(kern, params); }, __PRETTY_FUNCTION__);
	}
};
template <typename T>
inline unsigned long long to_bits(T v) {
	static_assert(sizeof(T) <= 8, "collective value wider than 64 bits");
	unsigned long long b = 0;
	memcpy(&b, &v, sizeof(T));
	return b;
}
template <typename T>
inline T from_bits(unsigned long long b) {
	T v;
	memcpy(&v, &b, sizeof(T));
	return v;
}
}  // namespace simt

#define threadIdx (simt::g_cur->tid)
extern uint3 blockIdx;
extern dim3 blockDim, gridDim;

// ------------------------------------------------------------------------------------------------ device intrinsics
static inline unsigned __float_as_uint(float f) { return simt::from_bits<unsigned>(simt::to_bits(f)); }
static inline float __uint_as_float(unsigned u) { return simt::from_bits<float>(simt::to_bits(u)); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline void __threadfence() {}
static inline void __syncthreads() { simt::cta_barrier(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
	unsigned long long out[32];
	simt::warp_exchange(mask, simt::OP_SYNCWARP, 0, out);
}
template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src, int = 32) {
	unsigned long long out[32];
	const unsigned part = simt::warp_exchange(mask, simt::OP_SHFL, simt::to_bits(v), out);
	src &= 31;
	return ((part >> src) & 1u) ? simt::from_bits<T>(out[src]) : v;
}
template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int = 32) {
	unsigned long long out[32];
	const unsigned part = simt::warp_exchange(mask, simt::OP_SHFL_UP, simt::to_bits(v), out);
	const unsigned lane = simt::g_cur->lane;
	if (lane < delta) return v;
	return ((part >> (lane - delta)) & 1u) ? simt::from_bits<T>(out[lane - delta]) : v;
}
template <typename T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int = 32) {
	unsigned long long out[32];
	const unsigned part = simt::warp_exchange(mask, simt::OP_SHFL_DOWN, simt::to_bits(v), out);
	const unsigned lane = simt::g_cur->lane;
	if (lane + delta > 31) return v;
	return ((part >> (lane + delta)) & 1u) ? simt::from_bits<T>(out[lane + delta]) : v;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
	unsigned long long out[32];
	const unsigned part = simt::warp_exchange(mask, simt::OP_BALLOT, pred ? 1ull : 0ull, out);
	unsigned r = 0;
	for (int i = 0; i < 32; ++i)
		if (((part >> i) & 1u) && out[i]) r |= 1u << i;
	return r;
}
template <typename T>
static inline unsigned __match_any_sync(unsigned mask, T v) {
	unsigned long long out[32];
	const unsigned long long mine = simt::to_bits(v);
	const unsigned part = simt::warp_exchange(mask, simt::OP_MATCH, mine, out);
	unsigned r = 0;
	for (int i = 0; i < 32; ++i)
		if (((part >> i) & 1u) && out[i] == mine) r |= 1u << i;
	return r;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
	unsigned long long out[32];
	const unsigned part = simt::warp_exchange(mask, simt::OP_REDUCE, v, out);
	unsigned r = 0;
	for (int i = 0; i < 32; ++i)
		if ((part >> i) & 1u) r += (unsigned)out[i];
	return r;
}
static inline int __reduce_min_sync(unsigned mask, int v) {
	unsigned long long out[32];
	const unsigned part = simt::warp_exchange(mask, simt::OP_REDUCE, simt::to_bits(v), out);
	int r = v;
	for (int i = 0; i < 32; ++i)
		if ((part >> i) & 1u) r = std::min(r, simt::from_bits<int>(out[i]));
	return r;
}
static inline int __reduce_max_sync(unsigned mask, int v) {
	unsigned long long out[32];
	const unsigned part = simt::warp_exchange(mask, simt::OP_REDUCE, simt::to_bits(v), out);
	int r = v;
	for (int i = 0; i < 32; ++i)
		if ((part >> i) & 1u) r = std::max(r, simt::from_bits<int>(out[i]));
	return r;
}
template <typename T>
static inline T __ldcg(const T* p) { return *p; }
// Fibers are cooperative (one OS thread), so a plain read-modify-write is atomic.
template <typename T>
static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T>
static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <typename T>
static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T>
static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }

// CUDA's global-namespace integer / float min and max
#define SIMT_MINMAX(T)                                         \
	static inline T min(T a, T b) { return b < a ? b : a; } \
	static inline T max(T a, T b) { return a < b ? b : a; }
SIMT_MINMAX(int)
SIMT_MINMAX(unsigned)
SIMT_MINMAX(long)
SIMT_MINMAX(unsigned long)
SIMT_MINMAX(long long)
SIMT_MINMAX(unsigned long long)
SIMT_MINMAX(float)
SIMT_MINMAX(double)
#undef SIMT_MINMAX

// ------------------------------------------------------------------------------------------------ runtime API subset
enum cudaError_t { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
typedef struct CUstream_st* cudaStream_t;
typedef struct CUevent_st* cudaEvent_t;
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaDeviceProp { char name[256]; int major, minor; int multiProcessorCount; };
enum {
	cudaHostAllocPortable = 1, cudaHostAllocMapped = 2, cudaHostRegisterPortable = 1, cudaHostRegisterMapped = 2,
	cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1
};

cudaError_t simt_malloc(void** p, size_t bytes);
template <typename T>
static inline cudaError_t cudaMalloc(T** p, size_t bytes) { return simt_malloc((void**)p, bytes); }
cudaError_t cudaFree(void* p);
cudaError_t simt_host_alloc(void** p, size_t bytes);
template <typename T>
static inline cudaError_t cudaHostAlloc(T** p, size_t bytes, unsigned) { return simt_host_alloc((void**)p, bytes); }
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaHostRegister(void* p, size_t bytes, unsigned flags);
cudaError_t cudaHostUnregister(void* p);
cudaError_t simt_host_device_pointer(void** d, void* h);
template <typename T>
static inline cudaError_t cudaHostGetDevicePointer(T** d, void* h, unsigned) { return simt_host_device_pointer((void**)d, h); }
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
// what cudaPointerGetAttributes answers for host memory: page-locked (cudaHostAlloc / cudaHostRegister) or not
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
	void* d = nullptr;
	a->type = simt_host_device_pointer(&d, const_cast<void*>(p)) == cudaSuccess ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered;
	a->device = 0; a->devicePointer = d; a->hostPointer = const_cast<void*>(p);
	return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t);
cudaError_t cudaMemsetAsync(void* dst, int v, size_t n, cudaStream_t);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t*, unsigned);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t*, unsigned, int);
cudaError_t cudaStreamDestroy(cudaStream_t);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned);
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaEventCreate(cudaEvent_t*);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t*, unsigned);
cudaError_t cudaEventDestroy(cudaEvent_t);
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t);
cudaError_t cudaEventSynchronize(cudaEvent_t);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t);
cudaError_t cudaGetDeviceCount(int*);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp*, int);
cudaError_t cudaSetDevice(int);
cudaError_t cudaGetLastError();
const char* cudaGetErrorString(cudaError_t);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*);
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned);
cudaError_t cudaIpcCloseMemHandle(void*);
