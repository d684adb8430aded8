// TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory): the one cub entry point the C ABI layer calls.
#pragma once
#include <cuda_runtime.h>

namespace cub {
struct DeviceScan {
	template <typename In, typename Out>
	static cudaError_t ExclusiveSum(void* tmp, size_t& tmp_bytes, const In* in, Out* out, int n, cudaStream_t = nullptr) {
		if (!tmp) { tmp_bytes = 1; return cudaSuccess; }
		Out acc = 0;
		for (int i = 0; i < n; ++i) { const Out v = (Out)in[i]; out[i] = acc; acc += v; }
		return cudaSuccess;
	}
};
}  // namespace cub
