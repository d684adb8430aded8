// TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory). cg::coalesced_threads() returns the set of threads that are
// converged at the call, which is implementation-defined; the interpreter runs lanes one at a time, so every thread is a
// coalesced group of its own — a legal outcome the kernels must be (and are) correct for.
#pragma once
#include <cuda_runtime.h>

namespace cooperative_groups {
struct coalesced_group {
	unsigned thread_rank() const { return 0; }
	unsigned size() const { return 1; }
	template <typename T>
	T shfl(T v, int) const { return v; }
};
inline coalesced_group coalesced_threads() { return coalesced_group(); }
}  // namespace cooperative_groups
