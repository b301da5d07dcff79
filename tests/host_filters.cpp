// host_filters.cpp -- TEST INFRASTRUCTURE: the fp32 filters of motionplanningtoolkit_b200/csrc/collide_device.cuh
// (obb_disjoint_f32 of the fused and round-1 kernels, the three-valued tri_filter_f32 of every kernel) compiled for the
// CPU.  tests/test_host_library.py copies the regions of the header between "[host-test:begin]" and "[host-test:end]"
// into collide_filters_extracted.inc next to this file's copy; only the stand-ins below are not the header's own text.
#include <cmath>
#include <cstdint>

#define __device__
#define __forceinline__ inline
struct float4 { float x, y, z, w; };

namespace mptcc {
namespace {
constexpr float kU = 5.9604645e-8f;  // 2^-24 (collide_device.cuh)
#include "collide_filters_extracted.inc"
}  // namespace
}  // namespace mptcc

extern "C" {
// a, b: four float4 per node (node32 layout); rt13: R (9), T (3), S
int host_obb_disjoint_f32(const float *a, const float *b, const float *rt13) {
	using namespace mptcc;
	RelF x;
	for (int k = 0; k < 9; ++k) x.R[k] = rt13[k];
	for (int k = 0; k < 3; ++k) x.T[k] = rt13[9 + k];
	x.S = rt13[12];
	const float4 *A = reinterpret_cast<const float4 *>(a), *B = reinterpret_cast<const float4 *>(b);
	float d2;
	return obb_disjoint_f32(A[0], A[1], A[2], A[3], B[0], B[1], B[2], B[3], x, d2) ? 1 : 0;
}
// p, q: three float4 per triangle (tri32 layout); returns TRI_SEPARATED 0 / TRI_UNCERTAIN 1 / TRI_INTERSECT 2
int host_tri_filter_f32(const float *p, const float *q, const float *rt13) {
	using namespace mptcc;
	RelF x;
	for (int k = 0; k < 9; ++k) x.R[k] = rt13[k];
	for (int k = 0; k < 3; ++k) x.T[k] = rt13[9 + k];
	x.S = rt13[12];
	const float4 *P = reinterpret_cast<const float4 *>(p), *Q = reinterpret_cast<const float4 *>(q);
	return tri_filter_f32(P[0], P[1], P[2], Q[0], Q[1], Q[2], x);
}
}
