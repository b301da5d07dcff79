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
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline float __uint_as_float(uint32_t u) { float f; __builtin_memcpy(&f, &u, 4); return f; }
// double -> float rounded towards -inf / +inf
static inline float __double2float_rd(double x) { const float f = (float)x; return (double)f <= x ? f : std::nextafterf(f, -INFINITY); }
static inline float __double2float_ru(double x) { const float f = (float)x; return (double)f >= x ? f : std::nextafterf(f, INFINITY); }
// round-to-nearest products and sums without contraction (this file is built with -ffp-contract=off)
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }

#include "device_layout.hpp"

// ---- warp-synchronous device functions on the CPU -------------------------------------------------------------------
// A converged warp is replayed lane by lane: every warp intrinsic is an "exchange site"; in pass n each lane runs from the
// start up to site n, leaves its operand there and stops (exception), so after the pass site n holds all 32 operands and
// pass n + 1 can read it.  The functions tested have warp-uniform control flow, i.e. every lane passes the same sites.
#include <array>
#include <cstring>
#include <vector>
namespace simt {
struct Stop {};
static int lane = 0, site = 0, ready = 0;
static std::vector<std::array<uint64_t, 32>> table;
static inline int exchange(uint64_t mine) {
	const int s = site++;
	if ((int)table.size() <= s) table.resize((size_t)s + 1);
	table[(size_t)s][(size_t)lane] = mine;
	if (s >= ready) throw Stop{};
	return s;
}
template <class F>
static void run_warp(F f) {  // f(lane) is called until every lane runs to its end
	table.clear();
	for (ready = 0;; ++ready) {
		bool done = true;
		for (lane = 0; lane < 32; ++lane) {
			site = 0;
			try {
				f(lane);
			} catch (const Stop &) {
				done = false;
			}
		}
		if (done) return;
	}
}
}  // namespace simt
static inline unsigned __ballot_sync(unsigned, bool pred) {
	const int s = simt::exchange(pred ? 1u : 0u);
	unsigned b = 0u;
	for (int l = 0; l < 32; ++l) b |= simt::table[(size_t)s][(size_t)l] ? 1u << l : 0u;
	return b;
}
static inline bool __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0u; }
template <class T>
static inline T __shfl_sync(unsigned, T v, int src) {
	uint64_t bits = 0;
	std::memcpy(&bits, &v, sizeof v);
	const int s = simt::exchange(bits);
	T out;
	std::memcpy(&out, &simt::table[(size_t)s][(size_t)(src & 31)], sizeof out);
	return out;
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }

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
// The candidate record of the np2 / np3 pipeline as the broadphase kernel builds it (make_candidate_record2): obj13 = the
// object's R (9), T (3) and its identity flag; m1 = the object's model (ctr 3, rinf, cmax1, esum), m2 = the agent's (ctr 3,
// r2, cmax1, esum).  Out: rec16 = R (9), Tc (3), Tinv (3), S; rel12 = FCL's fp64 relative transform; the path-certificate
// threshold.  Returns 1 if the pose passed the rigidity test.
int host_candidate_record2(const double *obj13, const double *m1v, const double *m2v, const double *pose, unsigned fmt, float *rec16,
                           double *rel12, float *cert) {
	using namespace mptcc;
	DevObject ob{};
	for (int k = 0; k < 9; ++k) ob.R[k] = obj13[k];
	for (int k = 0; k < 3; ++k) ob.T[k] = obj13[9 + k];
	ob.identity = obj13[12] != 0.0 ? 1u : 0u;
	ob.model = 0;
	DevModel m1{}, m2{};
	for (int k = 0; k < 3; ++k) {
		m1.ctr[k] = m1v[k];
		m2.ctr[k] = m2v[k];
	}
	m1.rinf = (float)m1v[3]; m1.cmax1 = m1v[4]; m1.esum = m1v[5];
	m2.r2 = (float)m2v[3]; m2.cmax1 = m2v[4]; m2.esum = m2v[5];
	KernelParams p{};
	p.env.models = &m1;
	p.pose_format = fmt;
	float4 rec[4];
	uint2 base;
	const bool rigid = make_candidate_record2(p, pose, ob, m2, rec, rel12, base, cert);
	__builtin_memcpy(rec16, rec, 64);
	return rigid ? 1 : 0;
}
// The record of the fused kernel / round-1 pipeline (make_candidate_record): rec13 = R (9), Tc (3), S.
void host_candidate_record(const double *obj13, const double *m1v, const double *m2v, const double *pose, unsigned fmt, float *rec13) {
	using namespace mptcc;
	DevObject ob{};
	for (int k = 0; k < 9; ++k) ob.R[k] = obj13[k];
	for (int k = 0; k < 3; ++k) ob.T[k] = obj13[9 + k];
	ob.identity = obj13[12] != 0.0 ? 1u : 0u;
	DevModel m1{}, m2{};
	for (int k = 0; k < 3; ++k) {
		m1.ctr[k] = m1v[k];
		m2.ctr[k] = m2v[k];
	}
	m1.rinf = (float)m1v[3];
	m2.r2 = (float)m2v[3];
	KernelParams p{};
	p.env.models = &m1;
	p.pose_format = fmt;
	float4 rec[4];
	make_candidate_record(p, pose, ob, m2, 0u, rec);
	const float *f = reinterpret_cast<const float *>(rec);
	for (int k = 0; k < 12; ++k) rec13[k] = f[k];
	rec13[12] = f[12];
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
	bool res[32];
	simt::run_warp([&](int lane) { res[lane] = mptcc::exact_tri_intersect_warp(rel, P, Q, lane); });
	for (int l = 1; l < 32; ++l)
		if (res[l] != res[0]) return -1;  // (a warp-wide answer)
	return res[0] ? 1 : 0;
}
// exact_path_reaches: does FCL's descent reach the leaf pair of triangles (prim1, prim2)?  node64 / anc: the model's
// exact-pass arrays (mptcc_bvh_get_exact_arrays), depth = levels per ancestor chain.
int host_exact_path_reaches(const double *rel, const double *node64_1, const uint32_t *anc1, unsigned depth1, const double *node64_2,
                            const uint32_t *anc2, unsigned depth2, unsigned prim1, unsigned prim2) {
	using namespace mptcc;
	KernelParams p{};
	p.env.node64 = node64_1;
	p.env.anc = anc1;
	p.agent.node64 = node64_2;
	p.agent.anc = anc2;
	DevModel m1{}, m2{};
	m1.depth = depth1;
	m2.depth = depth2;
	bool res[32];
	simt::run_warp([&](int lane) { res[lane] = exact_path_reaches(p, m1, m2, rel, prim1, prim2, lane); });
	for (int l = 1; l < 32; ++l)
		if (res[l] != res[0]) return -1;
	return res[0] ? 1 : 0;
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
