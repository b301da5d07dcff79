// peaks.cu -- measured non-tensor denominators for the three-way roofline of the collision path
// (t_roof = max(bytes_HBM / BW_HBM, bytes_L2 / BW_L2, flops / FP32)).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/peaks tools/peaks.cu && /tmp/peaks
//
// Prints ONE JSON object:
//   fp32_tflops        dense FFMA rate (8 independent chains per thread, all SMs)
//   l2_read_gbs        coalesced 128-bit loads (ld.global.cg) over a 48 MB buffer that stays L2-resident
//   gather64_l2_gbs    every lane fetches a random 64-byte record (two 256-bit loads, as the traversal does) from a
//                      48 MB table (L2-resident)
//   gather64_hbm_gbs   the same from a 1 GiB table (mostly HBM)
//   gather32_l2_gbs    random 32-byte records (one 256-bit load) from the 48 MB table
// Timing: CUDA events, 3 warm-up launches, best of 10.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                          \
	do {                                                                               \
		cudaError_t e__ = (x);                                                         \
		if (e__ != cudaSuccess) {                                                      \
			fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e__));                  \
			exit(1);                                                                   \
		}                                                                              \
	} while (0)

__global__ void __launch_bounds__(256) ffma_kernel(float *out, int iters, float a, float b) {
	float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int k = 0; k < 16; ++k) {
			x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
			x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

__global__ void __launch_bounds__(256) l2_read_kernel(const float4 *buf, size_t n_vec, int passes, float *out) {
	float acc = 0.f;
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (int p = 0; p < passes; ++p)
		for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
			float4 v;
			asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(buf + i));
			acc += v.x + v.y + v.z + v.w;
		}
	if (acc == 12345.678f) out[0] = acc;
}

__device__ __forceinline__ uint32_t xorshift(uint32_t &s) {
	s ^= s << 13;
	s ^= s >> 17;
	s ^= s << 5;
	return s;
}

// REC = 64: two 256-bit loads per record; REC = 32: one.  4 independent records in flight per lane.
template <int REC>
__global__ void __launch_bounds__(256) gather_kernel(const float4 *tab, uint32_t n_rec, int iters, float *out) {
	uint32_t s = 0x9E3779B9u * (blockIdx.x * blockDim.x + threadIdx.x + 1u);
	float acc = 0.f;
	for (int i = 0; i < iters; ++i) {
		float4 v[4][REC / 16];
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const uint32_t r = (uint32_t)(((unsigned long long)xorshift(s) * n_rec) >> 32);
			const float4 *p = tab + (size_t)r * (REC / 16);
#pragma unroll
			for (int h = 0; h < REC / 32; ++h)
				asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
				             : "=f"(v[k][2 * h].x), "=f"(v[k][2 * h].y), "=f"(v[k][2 * h].z), "=f"(v[k][2 * h].w),
				               "=f"(v[k][2 * h + 1].x), "=f"(v[k][2 * h + 1].y), "=f"(v[k][2 * h + 1].z), "=f"(v[k][2 * h + 1].w)
				             : "l"(p + 2 * h));
		}
#pragma unroll
		for (int k = 0; k < 4; ++k)
#pragma unroll
			for (int h = 0; h < REC / 16; ++h) acc += v[k][h].x + v[k][h].w;
	}
	if (acc == 12345.678f) out[0] = acc;
}

template <class F>
static float best_ms(F launch) {
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	for (int i = 0; i < 3; ++i) launch();
	CK(cudaDeviceSynchronize());
	float best = 1e30f;
	for (int i = 0; i < 10; ++i) {
		CK(cudaEventRecord(e0));
		launch();
		CK(cudaEventRecord(e1));
		CK(cudaEventSynchronize(e1));
		float ms;
		CK(cudaEventElapsedTime(&ms, e0, e1));
		if (ms < best) best = ms;
	}
	CK(cudaGetLastError());
	return best;
}

int main() {
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, 0));
	const int sms = prop.multiProcessorCount;
	float *out;
	CK(cudaMalloc(&out, sizeof(float) * 256 * sms * 16));

	// ---- FP32 ----
	const int iters = 2048, grid = sms * 8;
	const float ms_f = best_ms([&] { ffma_kernel<<<grid, 256>>>(out, iters, 1.0001f, 0.5f); });
	const double flops = 2.0 * 8 * 16 * (double)iters * 256.0 * grid;
	const double fp32_tflops = flops / (ms_f * 1e-3) / 1e12;

	// ---- L2-resident coalesced read ----
	const size_t small_bytes = 48ull << 20, big_bytes = 1ull << 30;
	float4 *tab;
	CK(cudaMalloc(&tab, big_bytes));
	CK(cudaMemset(tab, 0, big_bytes));
	const int passes = 20;
	const float ms_l2 = best_ms([&] { l2_read_kernel<<<sms * 8, 256>>>(tab, small_bytes / 16, passes, out); });
	const double l2_gbs = (double)small_bytes * passes / (ms_l2 * 1e-3) / 1e9;

	// ---- gathers ----
	const int git = 256;
	const int ggrid = sms * 8;
	const double recs = 4.0 * git * 256.0 * ggrid;
	const float ms_g64 = best_ms([&] { gather_kernel<64><<<ggrid, 256>>>(tab, (uint32_t)(small_bytes / 64), git, out); });
	const float ms_g64h = best_ms([&] { gather_kernel<64><<<ggrid, 256>>>(tab, (uint32_t)(big_bytes / 64), git, out); });
	const float ms_g32 = best_ms([&] { gather_kernel<32><<<ggrid, 256>>>(tab, (uint32_t)(small_bytes / 32), git, out); });
	const float ms_g32h = best_ms([&] { gather_kernel<32><<<ggrid, 256>>>(tab, (uint32_t)(big_bytes / 32), git, out); });

	printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp32_tflops\": %.2f, \"l2_read_gbs\": %.1f, "
	       "\"gather64_l2_gbs\": %.1f, \"gather64_hbm_gbs\": %.1f, \"gather32_l2_gbs\": %.1f, \"gather32_hbm_gbs\": %.1f, "
	       "\"gather64_l2_grecs\": %.2f, \"gather32_l2_grecs\": %.2f, "
	       "\"how\": \"tools/peaks.cu: FFMA 8 chains/thread; ld.global.cg.v4 over 48 MB x%d; random 64/32-byte records "
	       "(ld.global.nc.v8.f32, 4 in flight per lane) from a 48 MB (L2) / 1 GiB (HBM) table; CUDA events, best of 10\"}\n",
	       prop.name, sms, fp32_tflops, l2_gbs, recs * 64 / (ms_g64 * 1e-3) / 1e9, recs * 64 / (ms_g64h * 1e-3) / 1e9,
	       recs * 32 / (ms_g32 * 1e-3) / 1e9, recs * 32 / (ms_g32h * 1e-3) / 1e9, recs / (ms_g64 * 1e-3) / 1e9,
	       recs / (ms_g32 * 1e-3) / 1e9, passes);
	return 0;
}
