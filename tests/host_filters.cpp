// host_filters.cpp -- TEST INFRASTRUCTURE: device functions of motionplanningtoolkit_b200/csrc/collide_device.cuh compiled
// for the CPU: the fp32 filters (obb_disjoint_f32 of the fused and round-1 kernels, the three-valued tri_filter_f32 of every
// kernel) and the fp64 restatement of FCL's pose decoding and obbDisjoint (decode_pose, exact_obb_disjoint).  tests/test_host_library.py copies the regions of the header between "[host-test:begin]" and "[host-test:end]"
// into collide_filters_extracted.inc next to this file's copy; only the stand-ins below are not the header's own text.
#include <cmath>
#include <cstdint>

#define __device__
#define __forceinline__ inline
#define __noinline__
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
// double -> float rounded towards -inf / +inf
static inline float __double2float_rd(double x) { const float f = (float)x; return (double)f <= x ? f : std::nextafterf(f, -INFINITY); }
static inline float __double2float_ru(double x) { const float f = (float)x; return (double)f >= x ? f : std::nextafterf(f, INFINITY); }
// round-to-nearest products and sums without contraction (this file is built with -ffp-contract=off)
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }

#include "device_layout.hpp"

// exact_tri_intersect_warp ends in one ballot over the 17 lanes that each test one axis: the stand-in records the calling
// lane's predicate, host_exact_tri_intersect calls the function once per lane and forms the ballot itself
static int g_lane = 0;
static bool g_pred[32];
static inline unsigned __ballot_sync(unsigned, bool pred) {
	g_pred[g_lane] = pred;
	return 0u;
}

namespace mptcc {
namespace {
constexpr unsigned FULL = 0xffffffffu;
constexpr float kU = 5.9604645e-8f;  // 2^-24 (collide_device.cuh)
#include "collide_filters_extracted.inc"
}  // namespace
}  // namespace mptcc

extern "C" {
// pose record -> R (9), T (3), q (w x y z)
void host_decode_pose(const double *rec, unsigned fmt, double *R, double *T, double *q4) {
	mptcc::Quat q{1.0, 0.0, 0.0, 0.0};
	mptcc::decode_pose(rec, fmt, R, T, q, true);
	q4[0] = q.w; q4[1] = q.x; q4[2] = q.y; q4[3] = q.z;
}
// world AABB of a link (CollisionObject::computeAABB): model = aabb_min (3), aabb_max (3), aabb_center (3), aabb_radius
void host_pose_world_box(const double *rec, unsigned fmt, const double *model10, double *box6, float *amin3, float *amax3) {
	mptcc::DevModel m{};
	for (int k = 0; k < 3; ++k) {
		m.aabb_min[k] = model10[k];
		m.aabb_max[k] = model10[3 + k];
		m.aabb_center[k] = model10[6 + k];
	}
	m.aabb_radius = model10[9];
	double box[6 * 32];
	mptcc::pose_world_box(rec, fmt, m, box, amin3, amax3);
	for (int k = 0; k < 6; ++k) box6[k] = box[k * 32];
}
// relativeTransform(tf1 = object (R1 9, T1 3), tf2 = pose (R2 9, T2 3)) -> R (9), T (3)
void host_relative_transform(const double *R1, const double *T1, const double *R2, const double *T2, double *out12) {
	mptcc::DevObject ob{};
	for (int k = 0; k < 9; ++k) ob.R[k] = R1[k];
	for (int k = 0; k < 3; ++k) ob.T[k] = T1[k];
	for (int k = 0; k < 12; ++k) out12[k] = mptcc::rel_element(ob, R2, T2, k);
}
// Intersect::intersect_Triangle as the kernels evaluate it: rel = R (9), T (3); P, Q = 9 doubles each (model coordinates)
int host_exact_tri_intersect(const double *rel, const double *P, const double *Q) {
	unsigned ballot = 0u;
	for (g_lane = 0; g_lane < 32; ++g_lane) {
		(void)mptcc::exact_tri_intersect_warp(rel, P, Q, g_lane);
		ballot |= g_pred[g_lane] ? 1u << g_lane : 0u;
	}
	return ballot == 0u ? 1 : 0;
}
// rel = R (9), T (3); n1, n2 = node64 records: axis (9), To (3), extent (3), size
int host_exact_obb_disjoint(const double *rel, const double *n1, const double *n2) { return mptcc::exact_obb_disjoint(rel, n1, n2) ? 1 : 0; }
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
