// edge_kernels.cu -- SURVEY 8f N2: edge interpolation on the device.
//
// Omnidirectional::getPoses (/root/reference/agents/omnidirectional.hpp:202-247) turns an edge (start, end)
// into translation-only poses at spacing dt.  Shipping 48 bytes per edge instead of 96 bytes per pose cuts
// the host->device traffic by the number of poses per edge (10 .. 347 in the reference's configurations).
// The expansion restates the reference arithmetic exactly (fp64, no FMA), so the generated pose records are
// bit-identical to what the host would have sent.
#include <cuda_runtime.h>

#include <cstdint>

#include "device_layout.hpp"

namespace mptcc {
namespace {

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }

// one thread per generated pose-set; `offsets` is the exclusive prefix sum of the per-edge pose counts
__global__ void expand_omni_edges_kernel(const double *__restrict__ edges, const unsigned long long *__restrict__ offsets,
                                         unsigned long long n_edges, double dt, double *__restrict__ poses) {
	const unsigned long long total = offsets[n_edges];
	for (unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; t < total;
	     t += (unsigned long long)gridDim.x * blockDim.x) {
		// edge e with offsets[e] <= t < offsets[e+1]
		unsigned long long lo = 0, hi = n_edges;
		while (hi - lo > 1) {
			const unsigned long long mid = (lo + hi) >> 1;
			if (offsets[mid] <= t) lo = mid;
			else hi = mid;
		}
		const double *e = edges + 6 * lo;
		const unsigned long long i = t - offsets[lo];
		const double sx = e[0], sy = e[1], sz = e[2];
		const double dx = dadd(e[3], -sx), dy = dadd(e[4], -sy), dz = dadd(e[5], -sz);
		const double dist = sqrt(dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz)));
		const double ratio = dist / dt;
		const unsigned int iterations = ratio >= 1.0 ? (ratio < 4294967295.0 ? (unsigned int)ratio : 4294967295u) : 0u;
		double x, y, z;
		if (iterations < 1) {  // {start, end}
			x = i == 0 ? sx : e[3];
			y = i == 0 ? sy : e[4];
			z = i == 0 ? sz : e[5];
		} else if (i < iterations) {
			const double step = dt / dist;
			const double s = dmul(step, (double)i);
			x = dadd(sx, dmul(s, dx));
			y = dadd(sy, dmul(s, dy));
			z = dadd(sz, dmul(s, dz));
		} else {  // the extra end pose (only when iterations*dt < dist)
			x = e[3];
			y = e[4];
			z = e[5];
		}
		double *o = poses + 3 * t;  // a T3 record: fcl::Transform3f(Vec3f), identity rotation
		o[0] = x;
		o[1] = y;
		o[2] = z;
	}
}

// one thread per segment: safe iff no verdict bit of its pose-sets is set (Map3D::safeEdge)
__global__ void reduce_segments_kernel(const uint32_t *__restrict__ bits, const unsigned long long *__restrict__ offsets,
                                       unsigned long long n_segments, unsigned char *__restrict__ safe) {
	for (unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; s < n_segments;
	     s += (unsigned long long)gridDim.x * blockDim.x) {
		const unsigned long long b = offsets[s], e = offsets[s + 1];
		bool hit = false;
		for (unsigned long long i = b; i < e && !hit;) {
			const unsigned long long w = i >> 5;
			const unsigned lo = (unsigned)(i & 31ull);
			const unsigned long long end_w = ((w + 1) << 5) < e ? ((w + 1) << 5) : e;
			const unsigned len = (unsigned)(end_w - i);
			const uint32_t mask = (len >= 32u ? 0xffffffffu : ((1u << len) - 1u)) << lo;
			hit = (bits[w] & mask) != 0u;
			i = end_w;
		}
		safe[s] = hit ? 0 : 1;
	}
}

}  // namespace

int launch_expand_omni_edges(const double *d_edges, const unsigned long long *d_offsets, unsigned long long n_edges,
                             unsigned long long total_poses, double dt, double *d_poses, void *stream) {
	if (total_poses == 0) return 0;
	const int threads = 256;
	const unsigned long long want = (total_poses + threads - 1) / threads;
	const int blocks = (int)(want < 148ull * 16 ? want : 148ull * 16);
	expand_omni_edges_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(d_edges, d_offsets, n_edges, dt, d_poses);
	return (int)cudaGetLastError();
}

int launch_reduce_segments(const uint32_t *d_bits, const unsigned long long *d_offsets, unsigned long long n_segments,
                           unsigned char *d_safe, void *stream) {
	if (n_segments == 0) return 0;
	const int threads = 256;
	const unsigned long long want = (n_segments + threads - 1) / threads;
	const int blocks = (int)(want < 148ull * 16 ? want : 148ull * 16);
	reduce_segments_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(d_bits, d_offsets, n_segments, d_safe);
	return (int)cudaGetLastError();
}

}  // namespace mptcc
