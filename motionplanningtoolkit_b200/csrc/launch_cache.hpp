// launch_cache.hpp -- one-time per (device, kernel) launch preparation: the opt-in to more than 48 KB of dynamic
// shared memory and the occupancy query cost a few microseconds each, which is noticeable on planner-sized calls
// (one edge = ~10 pose-sets per call).
#pragma once

#include <cuda_runtime.h>

#include <mutex>
#include <vector>

namespace mptcc {

// Returns cudaSuccess and the resident CTAs per SM of `kernel` at (threads, smem); the first call per device does
// the cudaFuncSetAttribute / occupancy query, later calls are a table lookup.
inline cudaError_t prepare_kernel(const void *kernel, int threads, size_t smem, int *ctas_per_sm) {
	struct Entry {
		int dev;
		const void *kernel;
		int threads;
		size_t smem;
		int ctas_per_sm;
	};
	static std::mutex mu;
	static std::vector<Entry> table;
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	std::lock_guard<std::mutex> lock(mu);
	for (const Entry &t : table)
		if (t.dev == dev && t.kernel == kernel && t.threads == threads && t.smem == smem) {
			if (ctas_per_sm) *ctas_per_sm = t.ctas_per_sm;
			return cudaSuccess;
		}
	e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
#ifdef MPTCC_CARVEOUT
	e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, MPTCC_CARVEOUT);
	if (e != cudaSuccess) return e;
#endif
	int n = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem);
	if (e != cudaSuccess) return e;
	table.push_back(Entry{dev, kernel, threads, smem, n});
	if (ctas_per_sm) *ctas_per_sm = n;
	return cudaSuccess;
}

}  // namespace mptcc
