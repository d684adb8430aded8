// TEST INFRASTRUCTURE ONLY — scheduler and runtime half of the SIMT interpreter described in include/cuda_runtime.h.
// One OS thread; the threads of a CTA are fibers that the scheduler resumes round-robin, each running until it finishes or
// reaches a rendezvous (__syncthreads / warp collective) that is not complete yet. CTAs run one after the other.
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <map>

uint3 blockIdx;
dim3 blockDim, gridDim;

namespace simt {

Fiber* g_cur = nullptr;

namespace {

constexpr unsigned MAX_THREADS = 1024;
constexpr size_t STACK_BYTES = 256 * 1024;

extern "C" void simt_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
	pushq %rbp
	pushq %rbx
	pushq %r12
	pushq %r13
	pushq %r14
	pushq %r15
	movq %rsp, (%rdi)
	movq %rsi, %rsp
	popq %r15
	popq %r14
	popq %r13
	popq %r12
	popq %rbx
	popq %rbp
	ret
.size simt_switch,.-simt_switch
)");

struct Warp {
	unsigned alive = 0, arrived = 0, part = 0, mask = 0;
	int op = 0;
	bool draining = false;
	unsigned long long vals[32];
};

struct Cta {
	unsigned n = 0, alive = 0, bar_arrived = 0, bar_gen = 0;
	Fiber fibers[MAX_THREADS];
	Warp warps[MAX_THREADS / 32];
};

Cta g_cta;
// AXR_SIMT_ORDER = fwd (default) | rev | shuffle:<seed> — the order in which the scheduler visits the threads of a CTA and the
// CTAs of a grid. Every order is a legal execution of the same launch, so results that differ between orders mean the kernels
// depend on a schedule (a race, or an order assumption the hardware does not guarantee).
int g_order = -1;  // 0 fwd, 1 rev, 2 shuffle
unsigned long long g_seed = 1;
std::vector<unsigned> g_thread_order, g_cta_order;

void read_order() {
	if (g_order >= 0) return;
	const char* e = getenv("AXR_SIMT_ORDER");
	g_order = 0;
	if (e && !strcmp(e, "rev")) g_order = 1;
	else if (e && !strncmp(e, "shuffle", 7)) { g_order = 2; if (e[7] == ':') g_seed = strtoull(e + 8, nullptr, 10) * 2654435761ull + 1; }
}
void make_order(std::vector<unsigned>& v, size_t n, unsigned long long salt) {
	v.resize(n);
	for (size_t i = 0; i < n; ++i) v[i] = (unsigned)(g_order == 1 ? n - 1 - i : i);
	if (g_order == 2) {
		unsigned long long x = g_seed ^ (salt * 0x9E3779B97F4A7C15ull);
		for (size_t i = n; i > 1; --i) {  // Fisher-Yates with xorshift64*
			x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
			const size_t j = (size_t)((x * 2685821657736338717ull) >> 11) % i;
			std::swap(v[i - 1], v[j]);
		}
	}
}
char* g_stacks = nullptr;
void* g_sched_sp = nullptr;
const std::function<void()>* g_body = nullptr;
const char* g_kernel = "";
unsigned long long g_progress = 0;

[[noreturn]] void fatal(const char* what) {
	fprintf(stderr, "simt: %s\n  kernel %s\n  block (%u,%u,%u) of (%u,%u,%u), %u threads\n", what, g_kernel, blockIdx.x, blockIdx.y,
	        blockIdx.z, gridDim.x, gridDim.y, gridDim.z, g_cta.n);
	if (g_cur) fprintf(stderr, "  thread %u (warp %u lane %u)\n", g_cur->linear, g_cur->warp, g_cur->lane);
	for (unsigned w = 0; w * 32 < g_cta.n; ++w) {
		const Warp& W = g_cta.warps[w];
		fprintf(stderr, "  warp %u: alive %08x arrived %08x op %d draining %d\n", w, W.alive, W.arrived, W.op, (int)W.draining);
	}
	fprintf(stderr, "  barrier: arrived %u of %u alive\n", g_cta.bar_arrived, g_cta.alive);
	abort();
}

inline void yield() { simt_switch(&g_cur->sp, g_sched_sp); }

void release_barrier() {
	g_cta.bar_arrived = 0;
	g_cta.bar_gen++;
	g_progress++;
}

void fiber_entry() {
	(*g_body)();
	Fiber* f = g_cur;
	f->done = true;
	g_cta.warps[f->warp].alive &= ~(1u << f->lane);
	g_cta.alive--;
	g_progress++;
	// a thread that has exited no longer takes part in barriers (bar.sync counts the CTA's live threads on current hardware)
	if (g_cta.bar_arrived && g_cta.bar_arrived == g_cta.alive) release_barrier();
	yield();
	fatal("finished fiber resumed");
}

void run_cta(unsigned nthreads) {
	Cta& c = g_cta;
	c.n = c.alive = nthreads;
	c.bar_arrived = 0;
	for (unsigned w = 0; w * 32 < nthreads; ++w) {
		Warp& W = c.warps[w];
		const unsigned lanes = std::min(32u, nthreads - w * 32);
		W.alive = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1u);
		W.arrived = 0;
		W.draining = false;
	}
	for (unsigned i = 0; i < nthreads; ++i) {
		Fiber& f = c.fibers[i];
		f.linear = i;
		f.lane = i & 31;
		f.warp = i >> 5;
		f.tid.x = i % blockDim.x;
		f.tid.y = (i / blockDim.x) % blockDim.y;
		f.tid.z = i / (blockDim.x * blockDim.y);
		f.done = false;
		void** top = reinterpret_cast<void**>(g_stacks + (size_t)(i + 1) * STACK_BYTES);  // 16-byte aligned
		top[-1] = nullptr;                                  // where fiber_entry's caller's return address would be
		top[-2] = reinterpret_cast<void*>(&fiber_entry);   // popped by simt_switch's ret
		for (int k = 3; k <= 8; ++k) top[-k] = nullptr;     // rbp rbx r12 r13 r14 r15
		f.sp = top - 8;
	}
	unsigned remaining = nthreads;
	if (g_thread_order.size() != nthreads || g_order == 2)
		make_order(g_thread_order, nthreads, ((unsigned long long)blockIdx.z << 40) ^ ((unsigned long long)blockIdx.y << 20) ^ blockIdx.x);
	while (remaining) {
		const unsigned long long before = g_progress;
		for (unsigned k = 0; k < nthreads; ++k) {
			Fiber& f = c.fibers[g_thread_order[k]];
			if (f.done) continue;
			g_cur = &f;
			simt_switch(&g_sched_sp, f.sp);
			if (f.done) --remaining;
		}
		g_cur = nullptr;
		if (remaining && g_progress == before) fatal("deadlock: a full pass over the CTA's threads made no progress");
	}
}

}  // namespace

unsigned warp_exchange(unsigned mask, int op, unsigned long long v, unsigned long long out[32]) {
	Fiber* f = g_cur;
	if (!f) fatal("warp collective outside a kernel");
	Warp& W = g_cta.warps[f->warp];
	const unsigned bit = 1u << f->lane;
	if (!(mask & bit)) fatal("a lane executed a *_sync collective whose mask does not name it");
	while (W.draining) yield();  // the previous collective is still being read by slower lanes
	if (W.arrived == 0) { W.op = op; W.mask = mask; }
	else if (W.op != op || W.mask != mask) fatal("lanes of one warp met at different collectives (or with different masks)");
	W.vals[f->lane] = v;
	W.arrived |= bit;
	g_progress++;
	while (!W.draining && (W.arrived & mask & W.alive) != (mask & W.alive)) yield();
	if (!W.draining) { W.draining = true; W.part = W.arrived; }
	for (int i = 0; i < 32; ++i) out[i] = W.vals[i];
	const unsigned part = W.part;
	W.arrived &= ~bit;
	g_progress++;
	if (W.arrived == 0) W.draining = false;
	return part;
}

void cta_barrier() {
	if (!g_cur) fatal("__syncthreads outside a kernel");
	Cta& c = g_cta;
	const unsigned gen = c.bar_gen;
	c.bar_arrived++;
	g_progress++;
	if (c.bar_arrived == c.alive) { release_barrier(); return; }
	while (c.bar_gen == gen) yield();
}

void check_canaries(const char* when);

void run_grid(dim3 grid, dim3 block, const std::function<void()>& body, const char* name) {
	const unsigned long long nthreads = (unsigned long long)block.x * block.y * block.z;
	if (g_cur) fatal("nested kernel launch");
	g_kernel = name;
	gridDim = grid; blockDim = block;
	if (nthreads == 0 || nthreads > MAX_THREADS || grid.x == 0 || grid.y == 0 || grid.z == 0 || grid.y > 65535 || grid.z > 65535) {
		blockIdx = {0, 0, 0};
		g_cta.n = 0;
		fatal("invalid launch configuration");
	}
	if (!g_stacks) {
		g_stacks = static_cast<char*>(mmap(nullptr, MAX_THREADS * STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
		if (g_stacks == MAP_FAILED) { perror("simt: mmap"); abort(); }
	}
	g_body = &body;
	read_order();
	const size_t nctas = (size_t)grid.x * grid.y * grid.z;
	if (g_cta_order.size() != nctas || g_order == 2) make_order(g_cta_order, nctas, nctas);
	g_thread_order.clear();
	for (size_t k = 0; k < nctas; ++k) {
		const size_t id = g_cta_order[k];
		blockIdx = {(unsigned)(id % grid.x), (unsigned)(id / grid.x % grid.y), (unsigned)(id / ((size_t)grid.x * grid.y))};
		run_cta((unsigned)nthreads);
	}
	g_body = nullptr;
	check_canaries("after a kernel");
}

// ------------------------------------------------------------------------------------------------ memory
namespace {
// Built with -fsanitize=address (build.py sanitize="address") the allocations carry no zones of their own: AddressSanitizer's
// redzones then sit right behind the requested size and catch out-of-bounds READS by kernels as well, not only stores.
#if defined(__SANITIZE_ADDRESS__)
constexpr size_t GUARD = 0;
#else
constexpr size_t GUARD = 256;
#endif
constexpr unsigned char CANARY = 0xA5;
std::map<char*, size_t> g_dev;                 // user pointer -> bytes
std::map<char*, size_t> g_host;                // host ranges a device pointer can be asked for
}  // namespace

void check_canaries(const char* when) {
	for (auto& kv : g_dev) {
		const unsigned char* p = reinterpret_cast<unsigned char*>(kv.first);
		for (size_t i = 0; i < GUARD; ++i)
			if (p[-(long)GUARD + (long)i] != CANARY || p[kv.second + i] != CANARY) {
				fprintf(stderr, "simt: out-of-bounds store detected %s near device allocation %p (%zu bytes), %s it, last kernel %s\n", when,
				        (void*)p, kv.second, p[kv.second + i] != CANARY ? "after" : "before", g_kernel);
				abort();
			}
	}
}

}  // namespace simt

using namespace simt;

cudaError_t simt_malloc(void** p, size_t bytes) {
	char* base = static_cast<char*>(GUARD ? aligned_alloc(256, (bytes + 2 * GUARD + 255) / 256 * 256) : malloc(bytes ? bytes : 1));
	if (!base) return cudaErrorMemoryAllocation;
	memset(base, CANARY, GUARD);
	memset(base + GUARD, 0xCD, bytes);  // device memory is not zero-initialised
	memset(base + GUARD + bytes, CANARY, GUARD);
	g_dev[base + GUARD] = bytes;
	*p = base + GUARD;
	return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
	if (!p) return cudaSuccess;
	auto it = g_dev.find(static_cast<char*>(p));
	if (it == g_dev.end()) return cudaErrorInvalidValue;
	check_canaries("at cudaFree");
	free(it->first - GUARD);
	g_dev.erase(it);
	return cudaSuccess;
}
cudaError_t simt_host_alloc(void** p, size_t bytes) {
	char* q = static_cast<char*>(aligned_alloc(4096, (bytes + 4095) / 4096 * 4096));
	if (!q) return cudaErrorMemoryAllocation;
	g_host[q] = bytes;
	*p = q;
	return cudaSuccess;
}
cudaError_t cudaFreeHost(void* p) {
	auto it = g_host.find(static_cast<char*>(p));
	if (it == g_host.end()) return cudaErrorInvalidValue;
	g_host.erase(it);
	free(p);
	return cudaSuccess;
}
cudaError_t cudaHostRegister(void* p, size_t bytes, unsigned) {
	g_host[static_cast<char*>(p)] = bytes;
	return cudaSuccess;
}
cudaError_t cudaHostUnregister(void* p) {
	g_host.erase(static_cast<char*>(p));
	return cudaSuccess;
}
cudaError_t simt_host_device_pointer(void** d, void* h) {
	char* q = static_cast<char*>(h);
	auto it = g_host.upper_bound(q);
	if (it == g_host.begin()) return cudaErrorInvalidValue;
	--it;
	if (q >= it->first + it->second) return cudaErrorInvalidValue;
	*d = h;
	return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* dst, int v, size_t n, cudaStream_t) { memset(dst, v, n); return cudaSuccess; }

// Launches execute synchronously in issue order, which is one of the orders the streams and events of the C ABI layer allow
// (every wait is on an event recorded earlier), so streams and events carry no state.
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = reinterpret_cast<cudaStream_t>(new int(0)); return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned f, int) { return cudaStreamCreateWithFlags(s, f); }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete reinterpret_cast<int*>(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { check_canaries("at a stream synchronisation"); return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { if (lo) *lo = 0; if (hi) *hi = -1; return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { check_canaries("at a device synchronisation"); return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = reinterpret_cast<cudaEvent_t>(new int(0)); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete reinterpret_cast<int*>(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
	memset(p, 0, sizeof *p);
	snprintf(p->name, sizeof p->name, "SIMT interpreter (CPU, tests only)");
	p->major = 10; p->minor = 0; p->multiProcessorCount = 1;
	return cudaSuccess;
}
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory" : "invalid value"; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof *h); memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

// Marker the product-side loader (axiomr_b200/api.py: load_library) looks for, so that this build can never be picked up by
// bench.py, smoke() or a user by accident: it is refused unless the test harness asks for it explicitly.
extern "C" int axr_simt_interpreter_marker(void) { return 1; }
