// collide_device.cuh -- device code shared by the kernels of libmptcc (collide_kernel.cu, split_kernels.cu,
// self_kernel.cu): batched agent-mesh vs environment-mesh collision.
//
// Replaces, per pose-set, one MeshHandler::isInCollision (/root/reference/utilities/meshhandler.hpp:187-243):
//   broadphase  : world-AABB overlap of every (environment object, agent link) pair, fp64, the exact
//                 arithmetic of fcl::CollisionObject::computeAABB + AABB::overlap  [FCL-recalled C.5, C.6]
//   narrowphase : BVH-vs-BVH traversal.  The traversal itself is a CONSERVATIVE fp32 pass (it may only
//                 visit more pairs than FCL, never fewer); a leaf pair that the fp32 triangle test cannot
//                 prove separated is re-decided in fp64 with FCL's operation order (no FMA), and a
//                 positive pair additionally replays FCL's own root-to-leaf sequence of fp64 OBB tests
//                 [FCL-recalled C.7-C.9], so the verdict bit equals FCL's.
//
// Mapping.  Broadphase phases are lane-per-pose (pose decode, world AABB, stackless walk of the cluster tree)
// and lane-per-(pose, object) (exact AABB::overlap); they produce candidates (pose-set, object) with a
// 64-byte fp32 record of the relative transform.  The narrowphase (traverse_all) gives every lane its own
// candidate and its own depth-first stack (local memory); idle lanes take the next unstarted candidate or
// steal a pending sub-tree from a busy lane, warp-synchronously.  Leaf pairs are queued and tested 32 at a
// time.  First contact ends the pose-set.  The same code runs in the fused kernel (collide_kernel.cu: a warp
// owns `ppw` consecutive pose-sets), in the two kernels of the split pipeline (split_kernels.cu) and in the
// link-vs-link kernel (self_kernel.cu); a small policy type (FusedTrav / SplitTrav / SelfTrav) says where hits
// are recorded and where the exact pass finds a candidate's poses.
//
// Error bounds of the fp32 pass are derived in DESIGN.md ("fp32 filter bounds").
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "device_layout.hpp"

namespace mptcc {
namespace {


constexpr unsigned FULL = 0xffffffffu;
typedef unsigned long long u64;
// plan word y: [31:24] candidate, [23] children are environment nodes, [22] single (root) test, [21:0] agent node
constexpr unsigned kPlanEnv = 1u << 23, kPlanSingle = 1u << 22, kPlanIdxMask = (1u << 22) - 1u;
constexpr float kU = 5.9604645e-8f;  // 2^-24, fp32 unit round-off

// [host-test:begin]  (fp64: pose decode and FCL's obbDisjoint, compared bit for bit with the oracle on the CPU)
// ------------------------------------------------------------------------------------------------
// fp64 helpers with FCL's evaluation order and no FMA contraction
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dadd_rn(a, -b); }
// a0*b0 + a1*b1 + a2*b2, left to right
__device__ __forceinline__ double ddot(double a0, double a1, double a2, double b0, double b1, double b2) {
	return dadd(dadd(dmul(a0, b0), dmul(a1, b1)), dmul(a2, b2));
}
__device__ __forceinline__ double dabs(double a) { return fabs(a); }

struct Quat {
	double w, x, y, z;
};

// [FCL-recalled] Quaternion3f::operator*
__device__ __forceinline__ Quat qmul(const Quat &a, const Quat &b) {
	Quat r;
	r.w = dsub(dsub(dsub(dmul(a.w, b.w), dmul(a.x, b.x)), dmul(a.y, b.y)), dmul(a.z, b.z));
	r.x = dsub(dadd(dadd(dmul(a.w, b.x), dmul(a.x, b.w)), dmul(a.y, b.z)), dmul(a.z, b.y));
	r.y = dadd(dadd(dsub(dmul(a.w, b.y), dmul(a.x, b.z)), dmul(a.y, b.w)), dmul(a.z, b.x));
	r.z = dadd(dsub(dadd(dmul(a.w, b.z), dmul(a.x, b.y)), dmul(a.y, b.x)), dmul(a.z, b.w));
	return r;
}

// [FCL-recalled] Quaternion3f::fromRotation.  FCL branches on the trace and then on the largest
// diagonal entry; every branch is "root = sqrt(s); one component = 0.5 root; the other three =
// (sum or difference of two off-diagonal entries) * (0.5 / root)".  The operands are selected first so
// that the warp runs ONE sqrt and ONE division (both are long fp64 routines) instead of diverging.
__device__ __forceinline__ Quat quat_from_rotation(const double *R) {
	const double trace = dadd(dadd(R[0], R[4]), R[8]);
	const bool tpos = trace > 0.0;
	int i = 0;
	if (R[4] > R[0]) i = 1;
	if (R[8] > (i == 1 ? R[4] : R[0])) i = 2;
	// else-branch: s = ((Rii - Rjj) - Rkk) + 1 with (j,k) = (i+1, i+2) mod 3
	const double dii = i == 0 ? R[0] : (i == 1 ? R[4] : R[8]);
	const double djj = i == 0 ? R[4] : (i == 1 ? R[8] : R[0]);
	const double dkk = i == 0 ? R[8] : (i == 1 ? R[0] : R[4]);
	const double s = tpos ? dadd(trace, 1.0) : dadd(dsub(dsub(dii, djj), dkk), 1.0);
	double root = sqrt(s);
	const double half = dmul(0.5, root);
	root = 0.5 / root;
	// trace > 0 : x=(R21-R12) y=(R02-R20) z=(R10-R01)          (w = half)
	// i == 0    : w=(R21-R12) y=(R10+R01) z=(R20+R02)          (x = half)
	// i == 1    : w=(R02-R20) z=(R21+R12) x=(R01+R10)          (y = half)
	// i == 2    : w=(R10-R01) x=(R02+R20) y=(R12+R21)          (z = half)
	const double d21 = dsub(R[7], R[5]), d02 = dsub(R[2], R[6]), d10 = dsub(R[3], R[1]);
	const double s01 = dadd(R[3], R[1]), s02 = dadd(R[6], R[2]), s12 = dadd(R[7], R[5]);
	// a*b is commutative and R(j,i)+R(i,j) is symmetric in IEEE arithmetic, so the operand order of FCL's
	// additions does not matter; only which entries are combined does.
	const int sel = tpos ? 3 : i;
	const double A = sel == 3 || sel == 0 ? d21 : (sel == 1 ? d02 : d10);
	const double B = sel == 3 ? d02 : (sel == 0 ? s01 : (sel == 1 ? s12 : s02));
	const double C = sel == 3 ? d10 : (sel == 0 ? s02 : (sel == 1 ? s01 : s12));
	const double pa = dmul(A, root), pb = dmul(B, root), pc = dmul(C, root);
	Quat q;
	q.w = sel == 3 ? half : pa;
	q.x = sel == 3 ? pa : (sel == 0 ? half : (sel == 1 ? pc : pb));
	q.y = sel == 3 ? pb : (sel == 0 ? pb : (sel == 1 ? half : pc));
	q.z = sel == 3 ? pc : (sel == 0 ? pc : (sel == 1 ? pb : half));
	return q;
}

// [FCL-recalled] Quaternion3f::toRotation
__device__ __forceinline__ void quat_to_rotation(const Quat &q, double *R) {
	const double twoX = dmul(2.0, q.x), twoY = dmul(2.0, q.y), twoZ = dmul(2.0, q.z);
	const double twoWX = dmul(twoX, q.w), twoWY = dmul(twoY, q.w), twoWZ = dmul(twoZ, q.w);
	const double twoXX = dmul(twoX, q.x), twoXY = dmul(twoY, q.x), twoXZ = dmul(twoZ, q.x);
	const double twoYY = dmul(twoY, q.y), twoYZ = dmul(twoZ, q.y), twoZZ = dmul(twoZ, q.z);
	R[0] = dsub(1.0, dadd(twoYY, twoZZ));
	R[1] = dsub(twoXY, twoWZ);
	R[2] = dadd(twoXZ, twoWY);
	R[3] = dadd(twoXY, twoWZ);
	R[4] = dsub(1.0, dadd(twoXX, twoZZ));
	R[5] = dsub(twoYZ, twoWX);
	R[6] = dsub(twoXZ, twoWY);
	R[7] = dadd(twoYZ, twoWX);
	R[8] = dsub(1.0, dadd(twoXX, twoYY));
}

// Pose record -> R (9), T (3), q.  RT12: Transform3f(R,T) (q = fromRotation(R)); TQ7: Transform3f(q,T);
// T3: Transform3f(T) = identity rotation and identity quaternion (Omnidirectional::getPoses,
// /root/reference/agents/omnidirectional.hpp:224-242); XYZCS: x y z c s -> Transform3f(R,T) with the planar
// rotation the blimp and the snake fill in, R(0,0)=c R(1,0)=-s R(0,1)=s R(1,1)=c (/root/reference/agents/blimp.hpp:345-355,
// agents/snake_trailers.hpp:422-432), c and s evaluated by the caller's libm.
__device__ __forceinline__ void decode_pose(const double *rec, uint32_t fmt, double *R, double *T, Quat &q,
                                            bool want_q) {
	if (fmt == 2u || fmt == 3u) {
		const double c = fmt == 3u ? rec[3] : 1.0, s = fmt == 3u ? rec[4] : 0.0;
		R[0] = c;
		R[1] = s;
		R[2] = 0.0;
		R[3] = -s;
		R[4] = c;
		R[5] = 0.0;
		R[6] = 0.0;
		R[7] = 0.0;
		R[8] = 1.0;
		T[0] = rec[0];
		T[1] = rec[1];
		T[2] = rec[2];
		if (want_q) q = quat_from_rotation(R);  // (1,0,0,0) exactly for the identity, as Transform3f(T) stores
	} else if (fmt == 1u) {
		T[0] = rec[0];
		T[1] = rec[1];
		T[2] = rec[2];
		q.w = rec[3];
		q.x = rec[4];
		q.y = rec[5];
		q.z = rec[6];
		quat_to_rotation(q, R);
	} else {
#pragma unroll
		for (int k = 0; k < 9; ++k) R[k] = rec[k];
		T[0] = rec[9];
		T[1] = rec[10];
		T[2] = rec[11];
		if (want_q) q = quat_from_rotation(R);
	}
}

// ------------------------------------------------------------------------------------------------
// fp64 exact tests (rare path)
// ------------------------------------------------------------------------------------------------

// [FCL-recalled C.8] overlap(R0,T0,b1,b2) -> obbDisjoint; n1/n2 point at 16-double node64 records.
__device__ __noinline__ bool exact_obb_disjoint(const double *__restrict__ rel, const double *__restrict__ n1,
                                                const double *__restrict__ n2) {
	double A[9], Bx[9], c1[3], c2[3], a[3], b[3];
#pragma unroll
	for (int k = 0; k < 9; ++k) {
		A[k] = n1[k];
		Bx[k] = n2[k];
	}
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		c1[k] = n1[9 + k];
		c2[k] = n2[9 + k];
		a[k] = n1[12 + k];
		b[k] = n2[12 + k];
	}
	const double *R0 = rel, *T0 = rel + 9;
	// R0b2(r,c) = R0.row(r) . b2.axis[c]
	double M[9];
#pragma unroll
	for (int r = 0; r < 3; ++r)
#pragma unroll
		for (int c = 0; c < 3; ++c)
			M[3 * r + c] = ddot(R0[3 * r], R0[3 * r + 1], R0[3 * r + 2], Bx[3 * c], Bx[3 * c + 1], Bx[3 * c + 2]);
	// B(r,c) = R0b2.column(c) . b1.axis[r]
	double B[9];
#pragma unroll
	for (int r = 0; r < 3; ++r)
#pragma unroll
		for (int c = 0; c < 3; ++c) B[3 * r + c] = ddot(M[c], M[3 + c], M[6 + c], A[3 * r], A[3 * r + 1], A[3 * r + 2]);
	// Ttemp = R0 * b2.To + T0 - b1.To ; T[k] = Ttemp . b1.axis[k]
	double Tt[3], T[3];
#pragma unroll
	for (int r = 0; r < 3; ++r)
		Tt[r] = dsub(dadd(ddot(R0[3 * r], R0[3 * r + 1], R0[3 * r + 2], c2[0], c2[1], c2[2]), T0[r]), c1[r]);
#pragma unroll
	for (int k = 0; k < 3; ++k) T[k] = ddot(Tt[0], Tt[1], Tt[2], A[3 * k], A[3 * k + 1], A[3 * k + 2]);
	double Bf[9];
#pragma unroll
	for (int k = 0; k < 9; ++k) Bf[k] = dadd(dabs(B[k]), 1e-6);

	double t, s;
	// A0, B0, A1, A2, B1, B2
	t = dabs(T[0]);
	if (t > dadd(a[0], ddot(Bf[0], Bf[1], Bf[2], b[0], b[1], b[2]))) return true;
	s = ddot(B[0], B[3], B[6], T[0], T[1], T[2]);
	t = dabs(s);
	if (t > dadd(b[0], ddot(Bf[0], Bf[3], Bf[6], a[0], a[1], a[2]))) return true;
	t = dabs(T[1]);
	if (t > dadd(a[1], ddot(Bf[3], Bf[4], Bf[5], b[0], b[1], b[2]))) return true;
	t = dabs(T[2]);
	if (t > dadd(a[2], ddot(Bf[6], Bf[7], Bf[8], b[0], b[1], b[2]))) return true;
	s = ddot(B[1], B[4], B[7], T[0], T[1], T[2]);
	t = dabs(s);
	if (t > dadd(b[1], ddot(Bf[1], Bf[4], Bf[7], a[0], a[1], a[2]))) return true;
	s = ddot(B[2], B[5], B[8], T[0], T[1], T[2]);
	t = dabs(s);
	if (t > dadd(b[2], ddot(Bf[2], Bf[5], Bf[8], a[0], a[1], a[2]))) return true;

#define BB(r, c) B[3 * (r) + (c)]
#define BF(r, c) Bf[3 * (r) + (c)]
#define SUM4(p, q, r_, s_) dadd(dadd(dadd(p, q), r_), s_)
	// A0 x B0..B2
	s = dsub(dmul(T[2], BB(1, 0)), dmul(T[1], BB(2, 0)));
	t = dabs(s);
	if (t > SUM4(dmul(a[1], BF(2, 0)), dmul(a[2], BF(1, 0)), dmul(b[1], BF(0, 2)), dmul(b[2], BF(0, 1)))) return true;
	s = dsub(dmul(T[2], BB(1, 1)), dmul(T[1], BB(2, 1)));
	t = dabs(s);
	if (t > SUM4(dmul(a[1], BF(2, 1)), dmul(a[2], BF(1, 1)), dmul(b[0], BF(0, 2)), dmul(b[2], BF(0, 0)))) return true;
	s = dsub(dmul(T[2], BB(1, 2)), dmul(T[1], BB(2, 2)));
	t = dabs(s);
	if (t > SUM4(dmul(a[1], BF(2, 2)), dmul(a[2], BF(1, 2)), dmul(b[0], BF(0, 1)), dmul(b[1], BF(0, 0)))) return true;
	// A1 x B0..B2
	s = dsub(dmul(T[0], BB(2, 0)), dmul(T[2], BB(0, 0)));
	t = dabs(s);
	if (t > SUM4(dmul(a[0], BF(2, 0)), dmul(a[2], BF(0, 0)), dmul(b[1], BF(1, 2)), dmul(b[2], BF(1, 1)))) return true;
	s = dsub(dmul(T[0], BB(2, 1)), dmul(T[2], BB(0, 1)));
	t = dabs(s);
	if (t > SUM4(dmul(a[0], BF(2, 1)), dmul(a[2], BF(0, 1)), dmul(b[0], BF(1, 2)), dmul(b[2], BF(1, 0)))) return true;
	s = dsub(dmul(T[0], BB(2, 2)), dmul(T[2], BB(0, 2)));
	t = dabs(s);
	if (t > SUM4(dmul(a[0], BF(2, 2)), dmul(a[2], BF(0, 2)), dmul(b[0], BF(1, 1)), dmul(b[1], BF(1, 0)))) return true;
	// A2 x B0..B2
	s = dsub(dmul(T[1], BB(0, 0)), dmul(T[0], BB(1, 0)));
	t = dabs(s);
	if (t > SUM4(dmul(a[0], BF(1, 0)), dmul(a[1], BF(0, 0)), dmul(b[1], BF(2, 2)), dmul(b[2], BF(2, 1)))) return true;
	s = dsub(dmul(T[1], BB(0, 1)), dmul(T[0], BB(1, 1)));
	t = dabs(s);
	if (t > SUM4(dmul(a[0], BF(1, 1)), dmul(a[1], BF(0, 1)), dmul(b[0], BF(2, 2)), dmul(b[2], BF(2, 0)))) return true;
	s = dsub(dmul(T[1], BB(0, 2)), dmul(T[0], BB(1, 2)));
	t = dabs(s);
	if (t > SUM4(dmul(a[0], BF(1, 2)), dmul(a[1], BF(0, 2)), dmul(b[0], BF(2, 1)), dmul(b[1], BF(2, 0)))) return true;
#undef BB
#undef BF
#undef SUM4
	return false;
}
// [host-test:end]

// [host-test:begin]
struct D3 {
	double x, y, z;
};
__device__ __forceinline__ D3 d3sub(const D3 &a, const D3 &b) { return D3{dsub(a.x, b.x), dsub(a.y, b.y), dsub(a.z, b.z)}; }
__device__ __forceinline__ D3 d3cross(const D3 &l, const D3 &r) {
	return D3{dsub(dmul(l.y, r.z), dmul(l.z, r.y)), dsub(dmul(l.z, r.x), dmul(l.x, r.z)),
	          dsub(dmul(l.x, r.y), dmul(l.y, r.x))};
}
__device__ __forceinline__ double d3dot(const D3 &a, const D3 &b) { return ddot(a.x, a.y, a.z, b.x, b.y, b.z); }
__device__ __forceinline__ double dmin2(double a, double b) { return (b < a) ? b : a; }  // std::min
__device__ __forceinline__ double dmax2(double a, double b) { return (a < b) ? b : a; }  // std::max

// [FCL-recalled C.9] Intersect::intersect_Triangle(P1,P2,P3,Q1,Q2,Q3,R,T): the 17 axes are independent
// given the shared prefix, so lane k (k < 17) evaluates axis k; all 32 lanes must call.
// P, Q: 9 doubles each (model coordinates); rel: R (9), T (3).
__device__ __noinline__ bool exact_tri_intersect_warp(const double *__restrict__ rel, const double *__restrict__ P,
                                                      const double *__restrict__ Q, int lane) {
	const double *R = rel, *T = rel + 9;
	D3 P1{P[0], P[1], P[2]}, P2{P[3], P[4], P[5]}, P3{P[6], P[7], P[8]};
	D3 Qw[3];
#pragma unroll
	for (int i = 0; i < 3; ++i) {
		const double qx = Q[3 * i], qy = Q[3 * i + 1], qz = Q[3 * i + 2];
		Qw[i].x = dadd(ddot(R[0], R[1], R[2], qx, qy, qz), T[0]);
		Qw[i].y = dadd(ddot(R[3], R[4], R[5], qx, qy, qz), T[1]);
		Qw[i].z = dadd(ddot(R[6], R[7], R[8], qx, qy, qz), T[2]);
	}
	const D3 p1 = d3sub(P1, P1), p2 = d3sub(P2, P1), p3 = d3sub(P3, P1);
	const D3 q1 = d3sub(Qw[0], P1), q2 = d3sub(Qw[1], P1), q3 = d3sub(Qw[2], P1);
	const D3 e1 = d3sub(p2, p1), e2 = d3sub(p3, p2), e3 = d3sub(p1, p3);
	const D3 f1 = d3sub(q2, q1), f2 = d3sub(q3, q2), f3 = d3sub(q1, q3);
	const D3 n1 = d3cross(e1, e2), m1 = d3cross(f1, f2);

	D3 ax;
	if (lane == 0) ax = n1;
	else if (lane == 1) ax = m1;
	else {
		// lanes 2..10: e_a x f_b ; 11..13: e_a x n1 ; 14..16: f_a x m1
		D3 l, r;
		if (lane <= 10) {
			const int a = (lane - 2) / 3, b = (lane - 2) % 3;
			l = (a == 0) ? e1 : (a == 1) ? e2 : e3;
			r = (b == 0) ? f1 : (b == 1) ? f2 : f3;
		} else if (lane <= 13) {
			const int a = lane - 11;
			l = (a == 0) ? e1 : (a == 1) ? e2 : e3;
			r = n1;
		} else {
			const int a = (lane - 14) % 3;
			l = (a == 0) ? f1 : (a == 1) ? f2 : f3;
			r = m1;
		}
		ax = d3cross(l, r);
	}
	// project6
	const double a1 = d3dot(ax, p1), a2 = d3dot(ax, p2), a3 = d3dot(ax, p3);
	const double b1 = d3dot(ax, q1), b2 = d3dot(ax, q2), b3 = d3dot(ax, q3);
	const double mn1 = dmin2(a1, dmin2(a2, a3)), mx2 = dmax2(b1, dmax2(b2, b3));
	const double mx1 = dmax2(a1, dmax2(a2, a3)), mn2 = dmin2(b1, dmin2(b2, b3));
	const bool separated = (lane < 17) && ((mn1 > mx2) || (mn2 > mx1));
	return __ballot_sync(FULL, separated) == 0u;
}
// [host-test:end]

// [host-test:begin]
// Replays FCL's collisionRecurse descent from (root, root) to the leaf pair holding primitives
// (prim1, prim2) and evaluates every OBB test on the way in fp64 [FCL-recalled C.7, C.8].
// Returns true iff none of them reports "disjoint", i.e. iff FCL's traversal reaches that leaf pair.
//
// The root-to-leaf chain of each tree depends only on the primitive and is precomputed (ModelPack::anc):
// lane l fetches level l (and l+32) of both chains and the fp64 size of those nodes -- two dependent
// loads in total.  The FCL path is the merge of the two chains under the firstOverSecond rule; it is
// formed with shuffles and its pairs are tested 32 at a time.
__device__ __noinline__ bool exact_path_reaches(const KernelParams &p, const DevModel &m1, const DevModel &m2,
                                                const double *__restrict__ rel, uint32_t prim1, uint32_t prim2,
                                                int lane) {
	const uint32_t *__restrict__ anc1 = p.env.anc + m1.anc_base + (unsigned long long)prim1 * m1.depth;
	const uint32_t *__restrict__ anc2 = p.agent.anc + m2.anc_base + (unsigned long long)prim2 * m2.depth;
	const double *__restrict__ n64a = p.env.node64 + 16ull * m1.node_base;
	const double *__restrict__ n64b = p.agent.node64 + 16ull * m2.node_base;
	const uint32_t NONE = 0xffffffffu;
	const uint32_t na0 = (uint32_t)lane < m1.depth ? anc1[lane] : NONE;
	const uint32_t na1 = (uint32_t)lane + 32u < m1.depth ? anc1[lane + 32] : NONE;
	const uint32_t nb0 = (uint32_t)lane < m2.depth ? anc2[lane] : NONE;
	const uint32_t nb1 = (uint32_t)lane + 32u < m2.depth ? anc2[lane + 32] : NONE;
	const double sa0 = na0 != NONE ? n64a[16ull * na0 + 15] : 0.0;
	const double sa1 = na1 != NONE ? n64a[16ull * na1 + 15] : 0.0;
	const double sb0 = nb0 != NONE ? n64b[16ull * nb0 + 15] : 0.0;
	const double sb1 = nb1 != NONE ? n64b[16ull * nb1 + 15] : 0.0;
	// level of the leaf in each chain
	const int d1 = __popc(__ballot_sync(FULL, na0 != NONE)) + __popc(__ballot_sync(FULL, na1 != NONE)) - 1;
	const int d2 = __popc(__ballot_sync(FULL, nb0 != NONE)) + __popc(__ballot_sync(FULL, nb1 != NONE)) - 1;

	int i = 0, j = 0, k = 0;
	uint32_t my1 = 0, my2 = 0;
	bool have = false, disjoint_any = false;
	while (true) {
		const uint32_t ca = (i < 32) ? __shfl_sync(FULL, na0, i) : __shfl_sync(FULL, na1, i - 32);
		const uint32_t cb = (j < 32) ? __shfl_sync(FULL, nb0, j) : __shfl_sync(FULL, nb1, j - 32);
		const double sz1 = (i < 32) ? __shfl_sync(FULL, sa0, i) : __shfl_sync(FULL, sa1, i - 32);
		const double sz2 = (j < 32) ? __shfl_sync(FULL, sb0, j) : __shfl_sync(FULL, sb1, j - 32);
		if ((k & 31) == lane) {
			my1 = ca;
			my2 = cb;
			have = true;
		}
		++k;
		const bool l1 = i == d1, l2 = j == d2;
		const bool last = l1 && l2;
		if ((k & 31) == 0 || last) {
			bool dj = false;
			if (have) dj = exact_obb_disjoint(rel, n64a + 16ull * my1, n64b + 16ull * my2);
			have = false;
			if (__any_sync(FULL, dj)) disjoint_any = true;
			if (disjoint_any || last) break;
		}
		if (l2 || (!l1 && (sz1 > sz2))) ++i;
		else ++j;
	}
	return !disjoint_any;
}
// [host-test:end]

// ------------------------------------------------------------------------------------------------
// fp32 conservative tests (hot path)
// ------------------------------------------------------------------------------------------------

// [host-test:begin]  (tests/host_filters.cpp compiles the marked regions of this header for the CPU)
struct RelF {
	float R[9];
	float T[3];
	float S;  // tri filter scale (see tri_filter_f32)
};
// [host-test:end]

// One 256-bit read-only load (sm_100: LDG.E.256): half a node.  The L1 data pipe spends one wavefront per
// 32-byte sector an instruction touches; the per-lane gathers of the traversal touch a different sector in
// every lane, so a node fetched as four 16-byte loads costs four wavefronts per lane and as two 32-byte
// loads two -- that pipe is what bounds the narrowphase kernel.
struct F8 {
	float4 a, b;
};
#define MPTCC_LD256(HINT)                                                                                       \
	asm("ld.global.nc" HINT ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                           \
	    : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) \
	    : "l"(p))
__device__ __forceinline__ F8 ldg256(const float4 *p) {
	F8 r;
	MPTCC_LD256("");
	return r;
}
// L2 eviction priorities (256-bit accesses only, sm_100): tree nodes are re-read by every warp and should outlive
// the candidate list, which is written once by the broadphase kernel and read once by the narrowphase kernel.
__device__ __forceinline__ F8 ldg256_evict_last(const float4 *p) {
	F8 r;
	MPTCC_LD256(".L2::evict_last");
	return r;
}
__device__ __forceinline__ F8 ldg256_evict_first(const float4 *p) {
	F8 r;
	MPTCC_LD256(".L2::evict_first");
	return r;
}
__device__ __forceinline__ void stg256_evict_first(float4 *p, const float4 a, const float4 b) {
	asm volatile("st.global.L2::evict_first.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z),
	             "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
	             : "memory");
}
__device__ __forceinline__ void stg256(float4 *p, const float4 a, const float4 b) {
	asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w),
	             "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
	             : "memory");
}
__device__ __forceinline__ void stg256(double *p, double a, double b, double c, double d) {
	asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
// A pose record (12 doubles RT12 / 7 doubles TQ7) into registers; 32-byte aligned RT12 records (96 B) take
// three 256-bit loads instead of twelve 64-bit ones.
__device__ __forceinline__ void load_pose_record(const double *rec, uint32_t fmt, bool aligned32, double *out) {
	if (fmt == 0u && aligned32) {
#pragma unroll
		for (int k = 0; k < 3; ++k)
			asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
			    : "=d"(out[4 * k]), "=d"(out[4 * k + 1]), "=d"(out[4 * k + 2]), "=d"(out[4 * k + 3])
			    : "l"(rec + 4 * k));
	} else if (fmt == 0u) {
#pragma unroll
		for (int k = 0; k < 12; ++k) out[k] = rec[k];
	} else {
		const unsigned len = pose_rec_len(fmt);
#pragma unroll
		for (int k = 0; k < 7; ++k) out[k] = (unsigned)k < len ? rec[k] : 0.0;
#pragma unroll
		for (int k = 7; k < 12; ++k) out[k] = 0.0;
	}
}

// [host-test:begin]
// true => the two OBBs are certainly disjoint (with a margin far above fp64 round-off).
// n1x = node of the environment model, n2x = node of the agent model; margins: DESIGN.md.
__device__ __forceinline__ bool obb_disjoint_f32(const float4 a0, const float4 a1, const float4 a2, const float4 a3,
                                                 const float4 b0, const float4 b1, const float4 b2, const float4 b3,
                                                 const RelF &x, float &center_d2) {
	// M = R * [b axes as columns]
	float M[9];
#pragma unroll
	for (int r = 0; r < 3; ++r) {
		M[3 * r + 0] = x.R[3 * r] * b0.x + x.R[3 * r + 1] * b0.y + x.R[3 * r + 2] * b0.z;
		M[3 * r + 1] = x.R[3 * r] * b1.x + x.R[3 * r + 1] * b1.y + x.R[3 * r + 2] * b1.z;
		M[3 * r + 2] = x.R[3 * r] * b2.x + x.R[3 * r + 1] * b2.y + x.R[3 * r + 2] * b2.z;
	}
	// B(r,c) = a_r . M(:,c)
	float B[9];
#pragma unroll
	for (int c = 0; c < 3; ++c) {
		B[0 + c] = a0.x * M[c] + a0.y * M[3 + c] + a0.z * M[6 + c];
		B[3 + c] = a1.x * M[c] + a1.y * M[3 + c] + a1.z * M[6 + c];
		B[6 + c] = a2.x * M[c] + a2.y * M[3 + c] + a2.z * M[6 + c];
	}
	// Tt = R*c2 + T - c1 ; T = A^T Tt
	const float c2x = b0.w, c2y = b1.w, c2z = b2.w;
	const float c1x = a0.w, c1y = a1.w, c1z = a2.w;
	const float tx = x.R[0] * c2x + x.R[1] * c2y + x.R[2] * c2z + (x.T[0] - c1x);
	const float ty = x.R[3] * c2x + x.R[4] * c2y + x.R[5] * c2z + (x.T[1] - c1y);
	const float tz = x.R[6] * c2x + x.R[7] * c2y + x.R[8] * c2z + (x.T[2] - c1z);
	center_d2 = tx * tx + ty * ty + tz * tz;  // squared distance of the box centres (descent order heuristic)
	const float T0 = tx * a0.x + ty * a0.y + tz * a0.z;
	const float T1 = tx * a1.x + ty * a1.y + tz * a1.z;
	const float T2 = tx * a2.x + ty * a2.y + tz * a2.z;
	// margin = 128 u (|T|_1 + |c1|_1 + |c2|_1)
	const float S1 = (fabsf(x.T[0]) + fabsf(x.T[1]) + fabsf(x.T[2])) + (fabsf(c1x) + fabsf(c1y) + fabsf(c1z)) + (fabsf(c2x) + fabsf(c2y) + fabsf(c2z));
	const float mg = (128.0f * kU) * S1;
	float F[9];
#pragma unroll
	for (int k = 0; k < 9; ++k) F[k] = fabsf(B[k]) + 2e-6f;
	const float ea0 = a3.x, ea1 = a3.y, ea2 = a3.z, eb0 = b3.x, eb1 = b3.y, eb2 = b3.z;

	bool dj = false;
	dj |= fabsf(T0) > ea0 + (F[0] * eb0 + F[1] * eb1 + F[2] * eb2) + mg;
	dj |= fabsf(T1) > ea1 + (F[3] * eb0 + F[4] * eb1 + F[5] * eb2) + mg;
	dj |= fabsf(T2) > ea2 + (F[6] * eb0 + F[7] * eb1 + F[8] * eb2) + mg;
	dj |= fabsf(T0 * B[0] + T1 * B[3] + T2 * B[6]) > eb0 + (F[0] * ea0 + F[3] * ea1 + F[6] * ea2) + mg;
	dj |= fabsf(T0 * B[1] + T1 * B[4] + T2 * B[7]) > eb1 + (F[1] * ea0 + F[4] * ea1 + F[7] * ea2) + mg;
	dj |= fabsf(T0 * B[2] + T1 * B[5] + T2 * B[8]) > eb2 + (F[2] * ea0 + F[5] * ea1 + F[8] * ea2) + mg;
#ifndef MPTCC_OBB_FACE_ONLY
	// A0 x B0..2
	dj |= fabsf(T2 * B[3] - T1 * B[6]) > ea1 * F[6] + ea2 * F[3] + eb1 * F[2] + eb2 * F[1] + mg;
	dj |= fabsf(T2 * B[4] - T1 * B[7]) > ea1 * F[7] + ea2 * F[4] + eb0 * F[2] + eb2 * F[0] + mg;
	dj |= fabsf(T2 * B[5] - T1 * B[8]) > ea1 * F[8] + ea2 * F[5] + eb0 * F[1] + eb1 * F[0] + mg;
	// A1 x B0..2
	dj |= fabsf(T0 * B[6] - T2 * B[0]) > ea0 * F[6] + ea2 * F[0] + eb1 * F[5] + eb2 * F[4] + mg;
	dj |= fabsf(T0 * B[7] - T2 * B[1]) > ea0 * F[7] + ea2 * F[1] + eb0 * F[5] + eb2 * F[3] + mg;
	dj |= fabsf(T0 * B[8] - T2 * B[2]) > ea0 * F[8] + ea2 * F[2] + eb0 * F[4] + eb1 * F[3] + mg;
	// A2 x B0..2
	dj |= fabsf(T1 * B[0] - T0 * B[3]) > ea0 * F[3] + ea1 * F[0] + eb1 * F[8] + eb2 * F[7] + mg;
	dj |= fabsf(T1 * B[1] - T0 * B[4]) > ea0 * F[4] + ea1 * F[1] + eb0 * F[8] + eb2 * F[6] + mg;
	dj |= fabsf(T1 * B[2] - T0 * B[5]) > ea0 * F[5] + ea1 * F[2] + eb0 * F[7] + eb1 * F[6] + mg;
#endif
	return dj;
}

struct F3 {
	float x, y, z;
};
__device__ __forceinline__ F3 f3sub(const F3 &a, const F3 &b) { return F3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ F3 f3cross(const F3 &l, const F3 &r) {
	return F3{l.y * r.z - l.z * r.y, l.z * r.x - l.x * r.z, l.x * r.y - l.y * r.x};
}
__device__ __forceinline__ float f3dot(const F3 &a, const F3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float f3maxabs(const F3 &a) { return fmaxf(fabsf(a.x), fmaxf(fabsf(a.y), fabsf(a.z))); }

// One SAT axis of the fp32 filter: the larger of the two one-sided gaps along ax (p1 is the origin).
// gap > eps: certainly separated; gap < -eps: the projections certainly overlap (also in FCL's fp64 test).
__device__ __forceinline__ float axis_gap(const F3 &ax, const F3 &p2, const F3 &p3, const F3 &q1, const F3 &q2,
                                          const F3 &q3) {
	const float a2 = f3dot(ax, p2), a3 = f3dot(ax, p3);
	const float b1 = f3dot(ax, q1), b2 = f3dot(ax, q2), b3 = f3dot(ax, q3);
	const float mnP = fminf(0.0f, fminf(a2, a3)), mxP = fmaxf(0.0f, fmaxf(a2, a3));
	const float mnQ = fminf(b1, fminf(b2, b3)), mxQ = fmaxf(b1, fmaxf(b2, b3));
	return fmaxf(mnP - mxQ, mnQ - mxP);
}

// fp32 triangle filter, three-valued:
//   TRI_SEPARATED : some axis separates by more than the error bound -> FCL's fp64 test finds that axis too
//   TRI_INTERSECT : on all 17 axes the projections overlap by more than the error bound -> FCL's fp64 test
//                   finds no separating axis and reports the contact (degenerate, near-zero axes never qualify:
//                   their projections are all ~0, inside the band)
//   TRI_UNCERTAIN : anything else (including NaN) -> re-decided in fp64
constexpr int TRI_SEPARATED = 0, TRI_UNCERTAIN = 1, TRI_INTERSECT = 2;

__device__ __forceinline__ int tri_filter_f32(const float4 P1, const float4 P2, const float4 P3, const float4 Q1,
                                              const float4 Q2, const float4 Q3, const RelF &x) {
	const F3 o{P1.x, P1.y, P1.z};
	const F3 p2 = f3sub(F3{P2.x, P2.y, P2.z}, o), p3 = f3sub(F3{P3.x, P3.y, P3.z}, o);
	F3 q1, q2, q3;
	q1.x = x.R[0] * Q1.x + x.R[1] * Q1.y + x.R[2] * Q1.z + x.T[0];
	q1.y = x.R[3] * Q1.x + x.R[4] * Q1.y + x.R[5] * Q1.z + x.T[1];
	q1.z = x.R[6] * Q1.x + x.R[7] * Q1.y + x.R[8] * Q1.z + x.T[2];
	q2.x = x.R[0] * Q2.x + x.R[1] * Q2.y + x.R[2] * Q2.z + x.T[0];
	q2.y = x.R[3] * Q2.x + x.R[4] * Q2.y + x.R[5] * Q2.z + x.T[1];
	q2.z = x.R[6] * Q2.x + x.R[7] * Q2.y + x.R[8] * Q2.z + x.T[2];
	q3.x = x.R[0] * Q3.x + x.R[1] * Q3.y + x.R[2] * Q3.z + x.T[0];
	q3.y = x.R[3] * Q3.x + x.R[4] * Q3.y + x.R[5] * Q3.z + x.T[1];
	q3.z = x.R[6] * Q3.x + x.R[7] * Q3.y + x.R[8] * Q3.z + x.T[2];
	q1 = f3sub(q1, o);
	q2 = f3sub(q2, o);
	q3 = f3sub(q3, o);
	const float D = fmaxf(fmaxf(f3maxabs(p2), f3maxabs(p3)), fmaxf(f3maxabs(q1), fmaxf(f3maxabs(q2), f3maxabs(q3))));
	// DESIGN.md "fp32 filter bounds": eps2 = u D^2 (160 S + 512 D), eps3 = u D^3 (832 S + 2688 D)
	const float uD2 = kU * D * D;
	const float eps2 = uD2 * (160.0f * x.S + 512.0f * D);
	const float eps3 = uD2 * D * (832.0f * x.S + 2688.0f * D);

	const F3 e1 = p2, e2 = f3sub(p3, p2), e3 = F3{-p3.x, -p3.y, -p3.z};
	const F3 f1 = f3sub(q2, q1), f2 = f3sub(q3, q2), f3 = f3sub(q1, q3);
	const F3 n1 = f3cross(e1, e2), m1 = f3cross(f1, f2);
	// largest gap over the degree-2 axes and over the degree-3 axes (NaN-safe: fmaxf drops NaN, so NaNs are
	// tracked separately through `ok`)
	float g2 = axis_gap(n1, p2, p3, q1, q2, q3);
	bool ok = g2 == g2;
#define AX2(u, v)                                                   \
	{                                                               \
		const float g = axis_gap(f3cross(u, v), p2, p3, q1, q2, q3); \
		ok = ok && (g == g);                                        \
		g2 = fmaxf(g2, g);                                          \
	}
	{
		const float g = axis_gap(m1, p2, p3, q1, q2, q3);
		ok = ok && (g == g);
		g2 = fmaxf(g2, g);
	}
	AX2(e1, f1) AX2(e1, f2) AX2(e1, f3) AX2(e2, f1) AX2(e2, f2) AX2(e2, f3) AX2(e3, f1) AX2(e3, f2) AX2(e3, f3)
	if (ok && g2 > eps2) return TRI_SEPARATED;
	float g3 = axis_gap(f3cross(e1, n1), p2, p3, q1, q2, q3);
	ok = ok && (g3 == g3);
#define AX3(u, v)                                                   \
	{                                                               \
		const float g = axis_gap(f3cross(u, v), p2, p3, q1, q2, q3); \
		ok = ok && (g == g);                                        \
		g3 = fmaxf(g3, g);                                          \
	}
	AX3(e2, n1) AX3(e3, n1) AX3(f1, m1) AX3(f2, m1) AX3(f3, m1)
#undef AX2
#undef AX3
	if (!ok || !(eps2 == eps2) || !(eps3 == eps3)) return TRI_UNCERTAIN;
	if (g3 > eps3) return TRI_SEPARATED;
	if (g2 < -eps2 && g3 < -eps3) return TRI_INTERSECT;
	return TRI_UNCERTAIN;
}
// [host-test:end]

// ------------------------------------------------------------------------------------------------
// warp-cooperative traversal of one (environment object, agent link, pose) candidate
// ------------------------------------------------------------------------------------------------
struct WarpMem {
	uint2 *leafq;        // [kLeafQueueCap] (env prim, agent prim | cand << 24)
	float4 *cand;        // [kCandCap][4] candidate records: R(9) T(3) S src node_base1 tri_base1
	uint32_t *cand_ids;  // [kCandCap]    pose slot (0..63) << 26 | object
	uint32_t *clq;       // [kClusterQueueCap] pose slot << 27 | cluster
	double *rel;         // [16] fp64 R (9), T (3) of the candidate being re-decided exactly
	double *box;         // [6][32] world AABB of each lane's pose (min xyz, max xyz), component-major
	unsigned char *donor;  // [32] work stealing: donor rank -> lane
	// split pipeline only (narrowphase kernel):
	uint32_t *cand_pose;   // [kCandCap] pose-set index of the candidate
	unsigned char *dead;   // [kCandCap] per pose slot of the batch: already in collision
};

// [host-test:begin]
// relativeTransform(tf1, tf2): R = R1^T R2, T = R1^T (T2 - T1)  [FCL-recalled C.7]; element `idx`
// (0..8 -> R row-major, 9..11 -> T).
__device__ __forceinline__ double rel_element(const DevObject &ob, const double *R2, const double *T2, int idx) {
	if (idx < 9) {
		const int r = idx / 3, cc = idx % 3;
		return ddot(ob.R[r], ob.R[3 + r], ob.R[6 + r], R2[cc], R2[3 + cc], R2[6 + cc]);
	}
	const int r = idx - 9;
	const double d0 = dsub(T2[0], ob.T[0]), d1 = dsub(T2[1], ob.T[1]), d2 = dsub(T2[2], ob.T[2]);
	return ddot(ob.R[r], ob.R[3 + r], ob.R[6 + r], d0, d1, d2);
}
// [host-test:end]

// Candidate records live in shared memory as 4 float4 (64 B).  Lanes fetch the SAME part k of DIFFERENT
// records at once; with a plain layout all even (odd) candidates would hit the same 4 banks.  Part k of
// candidate c is therefore stored at position (k + (c >> 1)) & 3 of its record: 8 consecutive candidates
// cover all 32 banks and an LDS.128 of 8 lanes is one conflict-free wavefront.
__device__ __forceinline__ unsigned rec_at(unsigned cand, unsigned k) { return 4u * cand + ((k + (cand >> 1)) & 3u); }

// Hit bookkeeping of the fused kernel: one bit per pose slot of the warp's chunk, in registers.
struct FusedTrav {
	u64 hitmask, validmask;
	const double *link_poses;  // pose record of slot s of the current link: link_poses + s * stride
	size_t stride;
	__device__ __forceinline__ bool dead(unsigned slot, const WarpMem &) const { return (hitmask >> slot) & 1ull; }
	__device__ __forceinline__ bool all_done() const { return hitmask == validmask; }
	__device__ __forceinline__ void mark_hit(unsigned slot, unsigned, const WarpMem &, const KernelParams &, int) {
		hitmask |= 1ull << slot;
	}
	// every lane may bring one certified hit (slot >= 0); the mask stays warp-uniform
	__device__ __forceinline__ void mark_hits_per_lane(int slot, unsigned, const WarpMem &, int) {
		unsigned b = __ballot_sync(0xffffffffu, slot >= 0);
		while (b) {
			const int l = __ffs(b) - 1;
			b &= b - 1;
			hitmask |= 1ull << __shfl_sync(0xffffffffu, slot, l);
		}
	}
	__device__ __forceinline__ const double *pose_record(unsigned slot, unsigned, const WarpMem &) const {
		return link_poses + (size_t)slot * stride;
	}
	__device__ __forceinline__ unsigned object(unsigned cand, const WarpMem &wm) const { return wm.cand_ids[cand] & 0x3ffffffu; }
	// relativeTransform(object, pose) in fp64 -> wm.rel; lanes 0..11 compute one element each
	__device__ __forceinline__ const double *exact_rel(const KernelParams &p, const WarpMem &wm, const DevObject &ob,
	                                                   unsigned slot, unsigned cand, int lane) const {
		__syncwarp();
		if (lane < 12) {
			double R2[9], T2[3];
			Quat q;
			decode_pose(pose_record(slot, cand, wm), p.pose_format, R2, T2, q, false);
			wm.rel[lane] = rel_element(ob, R2, T2, lane);
		}
		__syncwarp();
		return wm.rel;
	}
};

// Hit bookkeeping of the narrowphase kernel (single-link agents): flags in shared memory per pose slot of
// the batch, verdict bits set with atomicOr (the buffer is zeroed before the pipeline runs).
struct SplitTrav {
	const double *poses;
	unsigned rec_len;
	uint32_t *verdict_bits;
	unsigned batch_first;  // list index of candidate 0 of the current batch
	__device__ __forceinline__ bool dead(unsigned slot, const WarpMem &wm) const { return wm.dead[slot] != 0; }
	__device__ __forceinline__ bool all_done() const { return false; }
	__device__ __forceinline__ void mark_hit(unsigned slot, unsigned cand, const WarpMem &wm, const KernelParams &, int lane) {
		__syncwarp();
		if (lane == 0) {
			wm.dead[slot] = 1;
			const uint32_t pose = wm.cand_pose[cand];
			atomicOr(verdict_bits + (pose >> 5), 1u << (pose & 31u));
		}
		__syncwarp();
	}
	__device__ __forceinline__ void mark_hits_per_lane(int slot, unsigned cand, const WarpMem &wm, int) {
		__syncwarp();
		if (slot >= 0) {
			wm.dead[slot] = 1;
			const uint32_t pose = wm.cand_pose[cand];
			atomicOr(verdict_bits + (pose >> 5), 1u << (pose & 31u));
		}
		__syncwarp();
	}
	__device__ __forceinline__ const double *pose_record(unsigned, unsigned cand, const WarpMem &wm) const {
		return poses + (size_t)wm.cand_pose[cand] * rec_len;
	}
	__device__ __forceinline__ unsigned object(unsigned cand, const WarpMem &wm) const { return wm.cand_ids[cand]; }
	// the broadphase kernel stored the fp64 relative transform next to the fp32 record
	__device__ __forceinline__ const double *exact_rel(const KernelParams &p, const WarpMem &, const DevObject &, unsigned,
	                                                   unsigned cand, int) const {
		return p.cand_rel + 12ull * (batch_first + cand);
	}
};

// [host-test:begin]
// Phase A for one lane: pose record -> world AABB of the link exactly as fcl::CollisionObject::computeAABB
// [FCL-recalled C.5] (identity quaternion: translated tight box; else the sphere box around
// q.transform(center) + T), stored component-major in `box` (stride 32 doubles) plus the outward-rounded
// fp32 copy used by the broadphase tree walk.
__device__ __forceinline__ void pose_world_box(const double *rec, uint32_t fmt, const DevModel &m2, double *box,
                                               float *amin, float *amax, double *rt_out = nullptr) {
	double R[9], T[3], bmin[3], bmax[3];
	Quat q;
	decode_pose(rec, fmt, R, T, q, true);
	if (rt_out != nullptr) {  // the decoded R (9), T (3) of the pose, component-major with stride 32 (one column per lane)
#pragma unroll
		for (int k = 0; k < 9; ++k) rt_out[k * 32] = R[k];
#pragma unroll
		for (int k = 0; k < 3; ++k) rt_out[(9 + k) * 32] = T[k];
	}
	if (q.w == 1.0 && q.x == 0.0 && q.y == 0.0 && q.z == 0.0) {
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			bmin[k] = dadd(m2.aabb_min[k], T[k]);
			bmax[k] = dadd(m2.aabb_max[k], T[k]);
		}
	} else {
		// Transform3f::transform(c) = q.transform(c) + T, q.transform(c) = (q (0,c) q*).xyz
		const Quat c{0.0, m2.aabb_center[0], m2.aabb_center[1], m2.aabb_center[2]};
		const Quat qc{q.w, -q.x, -q.y, -q.z};
		const Quat r = qmul(qmul(q, c), qc);
		const double cw[3] = {dadd(r.x, T[0]), dadd(r.y, T[1]), dadd(r.z, T[2])};
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			bmin[k] = dsub(cw[k], m2.aabb_radius);
			bmax[k] = dadd(cw[k], m2.aabb_radius);
		}
	}
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		box[k * 32] = bmin[k];
		box[(3 + k) * 32] = bmax[k];
		amin[k] = __double2float_rd(bmin[k]);
		amax[k] = __double2float_ru(bmax[k]);
	}
}
// [host-test:end]

// [host-test:begin]
// Phase C2 for one candidate: fp64 relativeTransform of (object, pose) -> fp32 record in the recentred
// frames: R (9), T (3), S, pose slot, node_base and tri_base of the object's model.
__device__ __forceinline__ void make_candidate_record(const KernelParams &p, const double *pose_rec, const DevObject &ob,
                                                      const DevModel &m2, unsigned slot, float4 *rec,
                                                      double *rel64 = nullptr) {
	const DevModel &m1 = p.env.models[ob.model];
	double R2[9], T2[3], rl[12];
	Quat q;
	decode_pose(pose_rec, p.pose_format, R2, T2, q, false);
#pragma unroll
	for (int k = 0; k < 12; ++k) rl[k] = rel_element(ob, R2, T2, k);
	if (rel64 != nullptr) {
#pragma unroll
		for (int k = 0; k < 12; ++k) rel64[k] = rl[k];  // FCL's R (9), T (3) verbatim, for the exact pass
	}
	// recentred frames: x1 - C1 = R (x2 - C2) + (R C2 + T - C1)
	float tf[3], tinf = 0.f;
#pragma unroll
	for (int r = 0; r < 3; ++r) {
		const double tc = rl[3 * r] * m2.ctr[0] + rl[3 * r + 1] * m2.ctr[1] + rl[3 * r + 2] * m2.ctr[2] + rl[9 + r] -
		                  m1.ctr[r];
		tf[r] = (float)tc;
		tinf = fmaxf(tinf, fabsf(tf[r]));
	}
	// S = max(S_q, S_p) + S_p, S_q = 5.01 r2(agent) + 4 |Tc|_inf, S_p = rinf(env)
	const float S = fmaxf(5.01f * m2.r2 + 4.0f * tinf * 1.0000002f, m1.rinf) + m1.rinf;
	rec[0] = make_float4((float)rl[0], (float)rl[1], (float)rl[2], (float)rl[3]);
	rec[1] = make_float4((float)rl[4], (float)rl[5], (float)rl[6], (float)rl[7]);
	rec[2] = make_float4((float)rl[8], tf[0], tf[1], tf[2]);
	rec[3] = make_float4(S, __uint_as_float(slot), __uint_as_float(m1.node_base), __uint_as_float(m1.tri_base));
}

// Candidate record of the second-generation narrowphase (np2_kernels.cu): R (9), Tc (3), Tinv (3), S, where
// x1 - C1 = R (x2 - C2) + Tc is the relative transform in the recentred frames and Tinv = -R^T Tc the translation
// of its inverse (fp64, rounded once).  `base` = node_base / tri_base of the object's model.  Returns false if R
// is not a rotation to within 1e-11 (then (R^T, Tinv) is not the inverse and the caller must route the launch to
// the fused kernel, which makes no such assumption).
// (R2, T2 = the decoded pose, see decode_pose)
__device__ __forceinline__ bool make_candidate_record2_rt(const KernelParams &p, const double *R2, const double *T2, const DevObject &ob,
                                                          const DevModel &m2, float4 *rec, double *rel64, uint2 &base,
                                                          float *cert = nullptr) {
	const DevModel &m1 = p.env.models[ob.model];
	double rl[12];
	if (ob.identity) {
		// R1 = I, T1 = 0 (every .inst file of the reference places the environment at the identity): FCL's
		// R1^T R2 and R1^T (T2 - T1) are 1 x + 0 y + 0 z, i.e. x itself for finite entries (up to the sign of a zero,
		// which no comparison sees); entries that are not finite make the launch fall back below, as they do with the
		// general formula (0 * inf = NaN there, inf here: both fail the finite test).
#pragma unroll
		for (int k = 0; k < 9; ++k) rl[k] = R2[k];
#pragma unroll
		for (int k = 0; k < 3; ++k) rl[9 + k] = T2[k];
	} else {
#pragma unroll
		for (int k = 0; k < 12; ++k) rl[k] = rel_element(ob, R2, T2, k);
	}
	if (rel64 != nullptr) {
#pragma unroll
		for (int k = 0; k < 12; ++k) rel64[k] = rl[k];  // FCL's R (9), T (3) (verbatim unless the object sits at the identity)
	}
	double tc[3];
	float tinf = 0.f;
#pragma unroll
	for (int r = 0; r < 3; ++r) {
		tc[r] = rl[3 * r] * m2.ctr[0] + rl[3 * r + 1] * m2.ctr[1] + rl[3 * r + 2] * m2.ctr[2] + rl[9 + r] - m1.ctr[r];
		tinf = fmaxf(tinf, fabsf((float)tc[r]));
	}
	double ti[3], defect = 0.0;
#pragma unroll
	for (int c = 0; c < 3; ++c) {
		ti[c] = -(rl[c] * tc[0] + rl[3 + c] * tc[1] + rl[6 + c] * tc[2]);
#pragma unroll
		for (int d = c; d < 3; ++d) {
			const double g = rl[c] * rl[d] + rl[3 + c] * rl[3 + d] + rl[6 + c] * rl[6 + d] - (c == d ? 1.0 : 0.0);
			defect = fmax(defect, fabs(g));  // (fmax drops NaN: the finite test below catches those)
		}
	}
	bool finite = true;
#pragma unroll
	for (int k = 0; k < 12; ++k) finite = finite && (fabs(rl[k]) <= 1.7e308);
	const float S = fmaxf(5.01f * m2.r2 + 4.0f * tinf * 1.0000002f, m1.rinf) + m1.rinf;
	rec[0] = make_float4((float)rl[0], (float)rl[1], (float)rl[2], (float)rl[3]);
	rec[1] = make_float4((float)rl[4], (float)rl[5], (float)rl[6], (float)rl[7]);
	rec[2] = make_float4((float)rl[8], (float)tc[0], (float)tc[1], (float)tc[2]);
	rec[3] = make_float4((float)ti[0], (float)ti[1], (float)ti[2], S);
	base = make_uint2(m1.node_base, m1.tri_base);
	if (cert != nullptr) {
		// Path certificate threshold: 1e-5 x an upper bound of the magnitudes entering FCL's OBB
		// tests on the path to any leaf pair of this candidate -- |T|_1 of the exact transform is bounded through the
		// recentred translation -- rounded up to fp32.  NaN compares false, i.e. "replay".
		const double mag = (fabs(tc[0]) + fabs(tc[1]) + fabs(tc[2])) * (1.0 + 4e-6) + fabs(m1.ctr[0]) + fabs(m1.ctr[1]) + fabs(m1.ctr[2]) +
		                   1.75 * (fabs(m2.ctr[0]) + fabs(m2.ctr[1]) + fabs(m2.ctr[2])) + m1.cmax1 + m2.cmax1 + m1.esum + m2.esum;
		*cert = __double2float_ru(1e-5 * mag * (1.0 + 1e-12));
	}
	return finite && defect <= 1e-11;
}
__device__ __forceinline__ bool make_candidate_record2(const KernelParams &p, const double *pose_rec, const DevObject &ob,
                                                       const DevModel &m2, float4 *rec, double *rel64, uint2 &base,
                                                       float *cert = nullptr) {
	double R2[9], T2[3];
	Quat q;
	decode_pose(pose_rec, p.pose_format, R2, T2, q, false);
	return make_candidate_record2_rt(p, R2, T2, ob, m2, rec, rel64, base, cert);
}
// [host-test:end]

// Pose box -> the integer grid of the broadphase tree (device_layout.hpp), rounded outward by one extra cell
// (the fp32 evaluation of (x - origin) * inv_cell is off by < 0.02 cells).  Non-finite boxes keep the old
// behaviour of the float test: NaN "overlaps" everything, +-inf saturates.
__device__ __forceinline__ void quantize_pose_box(const KernelParams &p, const float *amin, const float *amax, int *qmin,
                                                  int *qmax) {
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		const float lo = (amin[k] - p.bp_origin[k]) * p.bp_inv_cell[k], hi = (amax[k] - p.bp_origin[k]) * p.bp_inv_cell[k];
		qmin[k] = (int)floorf(fminf(fmaxf(lo, -2.0f), 65538.0f)) - 1;  // NaN -> -3
		qmax[k] = (int)ceilf(fmaxf(fminf(hi, 65538.0f), -2.0f)) + 1;   // NaN -> 65539
	}
}

// One step of the lane-per-pose walk of the threaded broadphase tree (one 16-byte node per visit); returns
// true and the cluster when the lane's quantized box touches a leaf.
__device__ __forceinline__ bool bp_tree_step(const KernelParams &p, unsigned &node, const int *qmin, const int *qmax,
                                             unsigned &cluster) {
	const uint4 nd = __ldg(p.bp_nodes + node);
	const int lx = (int)(nd.x & 0xffffu), ly = (int)(nd.x >> 16), lz = (int)(nd.y & 0xffffu);
	const int hx = (int)(nd.y >> 16), hy = (int)(nd.z & 0xffffu), hz = (int)(nd.z >> 16);
	const bool ov = !(lx > qmax[0] || ly > qmax[1] || lz > qmax[2] || hx < qmin[0] || hy < qmin[1] || hz < qmin[2]);
	const bool leaf = (nd.w >> 31) != 0u;
	cluster = nd.w & 0x7fffffffu;
	node = (leaf || ov) ? node + 1u : nd.w;  // (a leaf's escape is the next node in pre-order)
	return ov && leaf;
}

// FCL's exact closed-interval AABB::overlap(environment object, agent link) in fp64 [FCL-recalled C.6];
// `sl` = the lane slot whose pose box is in wm.box.
__device__ __forceinline__ bool exact_aabb_overlap(const KernelParams &p, const WarpMem &wm, unsigned eobj, unsigned sl) {
	const double *bb = p.obj_aabb + eobj;
	const unsigned st = p.n_objects_pad;
	const double e0 = bb[0], e1 = bb[st], e2 = bb[2 * st], e3 = bb[3 * st], e4 = bb[4 * st], e5 = bb[5 * st];
	const bool out = (e0 > wm.box[3 * 32 + sl]) | (e1 > wm.box[4 * 32 + sl]) | (e2 > wm.box[5 * 32 + sl]) |
	                 (e3 < wm.box[0 * 32 + sl]) | (e4 < wm.box[1 * 32 + sl]) | (e5 < wm.box[2 * 32 + sl]);
	return !out && eobj < p.n_objects;  // (a NaN box "overlaps" the padding boxes too)
}

// Narrowphase of all `ncand` candidates of the warp's pose chunk.  Every lane runs FCL's depth-first
// descent for one candidate at a time on a private stack (local memory, ring buffer); a lane whose
// stack runs empty takes the next unstarted candidate, and once those are gone it steals the bottom
// (= largest pending subtree) entry of a lane that has at least two -- all warp-synchronously, with
// ballots and shuffles only.  Leaf pairs go to a shared queue and are tested 32 at a time.  Depth-first
// order per candidate is what lets a colliding pose stop after about as many tests as FCL needs.
//
// `Trav` says how hits are recorded and where a candidate's pose lives: FusedTrav (bit mask over the pose
// slots of the warp's chunk, fused kernel) or SplitTrav (shared-memory flags + atomicOr into the global
// verdict words, narrowphase kernel of the split pipeline).
template <bool COUNT, class Trav>
__device__ void traverse_all(const KernelParams &p, const WarpMem &wm, const DevModel &m2, int ncand, Trav &tv,
                             int lane, unsigned long long *ctr) {
	const float4 *__restrict__ nodes2 = p.agent.node32 + 4ull * m2.node_base;
	const float4 *__restrict__ tris2 = p.agent.tri32 + 3ull * m2.tri_base;
	const unsigned lt = (1u << lane) - 1u;
	constexpr unsigned M = kLaneStackCap - 1;

	uint2 stk[kLaneStackCap];
	unsigned top = 0, bot = 0;  // ring buffer: entries [bot, top) = pending right siblings
	uint2 cur = make_uint2(0u, 0u);
	bool has = false;           // `cur` is the pair this lane tests next (top of its logical stack)
	int qhead = 0, lq = 0;

	while (!tv.all_done()) {
		if (!has && top != bot) {
			cur = stk[--top & M];
			has = true;
		}
		// ---- 1. idle lanes start the next unstarted candidates --------------------------------------
		unsigned im = __ballot_sync(FULL, !has);
		if (im && qhead < ncand) {
			const int c = qhead + __popc(im & lt);
			if (!has && c < ncand) {
				cur = make_uint2(0u, kPlanSingle | ((unsigned)c << 24));  // root pair
				has = true;
			}
			qhead += __popc(im);
			im = __ballot_sync(FULL, !has);
		}
		// ---- 2. ... or steal the bottom entry (largest pending subtree) of a busy lane ------------------
		if (im && qhead >= ncand && !p.no_steal) {
			const unsigned dm = __ballot_sync(FULL, has && top != bot);
			if (dm) {
				const int n_thieves = __popc(im), n_donors = __popc(dm);
				const int drank = __popc(dm & lt);
				const bool donor = (dm >> lane) & 1u;
				if (donor) wm.donor[drank] = (unsigned char)lane;  // rank -> lane table: the r-th thief robs the r-th donor
				__syncwarp();
				const int r = __popc(im & lt);
				const bool thief = !has && r < n_donors;
				const int from = thief ? (int)wm.donor[r] : lane;
				const bool robbed = donor && drank < n_thieves;
				uint2 e = make_uint2(0u, 0u);
				if (robbed) e = p.steal_bottom ? stk[bot++ & M] : stk[--top & M];
				e.x = __shfl_sync(FULL, e.x, from);
				e.y = __shfl_sync(FULL, e.y, from);
				if (thief) {
					cur = e;
					has = true;
				}
			}
		}
		const bool idle = !has;
		const unsigned am = __ballot_sync(FULL, !idle);
		if (am == 0u && lq == 0) break;

		if (lq >= 32 || am == 0u) {  // (a node step adds at most 64 entries: capacity 32 + 64)
			// ---- 3. leaf phase: up to 32 triangle pairs -------------------------------------------------
			const int k = lq < 32 ? lq : 32;
			lq -= k;
			int verdict = TRI_SEPARATED, fast_slot = -1;
			uint2 pr = make_uint2(0u, 0u);
			if (lane < k) {
				pr = wm.leafq[lq + lane];
				const unsigned lc = pr.y >> 24;
				const float4 r0 = wm.cand[rec_at(lc, 0)], r1 = wm.cand[rec_at(lc, 1)], r2 = wm.cand[rec_at(lc, 2)],
				             r3 = wm.cand[rec_at(lc, 3)];
				if (!tv.dead(__float_as_uint(r3.y), wm)) {
					RelF x;
					x.R[0] = r0.x; x.R[1] = r0.y; x.R[2] = r0.z; x.R[3] = r0.w;
					x.R[4] = r1.x; x.R[5] = r1.y; x.R[6] = r1.z; x.R[7] = r1.w;
					x.R[8] = r2.x; x.T[0] = r2.y; x.T[1] = r2.z; x.T[2] = r2.w;
					x.S = r3.x;
					const float4 *t1 = p.env.tri32 + 3ull * (__float_as_uint(r3.w) + pr.x);
					const float4 *t2 = tris2 + 3ull * (pr.y & 0xffffffu);
					const float4 P1 = __ldg(t1), P2 = __ldg(t1 + 1), P3 = __ldg(t1 + 2);
					const float4 Q1 = __ldg(t2), Q2 = __ldg(t2 + 1), Q3 = __ldg(t2 + 2);
					verdict = tri_filter_f32(P1, P2, P3, Q1, Q2, Q3, x);
					if (verdict == TRI_INTERSECT && !p.always_replay) {
						// Certain contact: try the path certificate (below) right here, every lane for its own
						// pair, so the dependent loads of all pairs overlap instead of queueing in the
						// warp-serial loop.  |T|_1 of the exact transform is bounded from the recentred fp32
						// translation of the record (a larger magnitude only makes the certificate stricter).
						const DevModel &m1 = p.env.models[p.objects[tv.object(lc, wm)].model];
						const double t1n = (double)((fabsf(x.T[0]) + fabsf(x.T[1]) + fabsf(x.T[2])) * 1.000001f);
						const double mag = t1n + fabs(m1.ctr[0]) + fabs(m1.ctr[1]) + fabs(m1.ctr[2]) +
						                   1.75 * (fabs(m2.ctr[0]) + fabs(m2.ctr[1]) + fabs(m2.ctr[2])) + m1.cmax1 +
						                   m2.cmax1 + m1.esum + m2.esum;
						const double mu = fmin((double)p.env.mu_chain[__float_as_uint(r3.w) + pr.x],
						                       (double)p.agent.mu_chain[m2.tri_base + (pr.y & 0xffffffu)]);
						fast_slot = mu >= 1e-5 * mag ? (int)__float_as_uint(r3.y) : -1;
					}
				}
			}
			if (COUNT) ctr[CTR_LEAF] += (unsigned)k;
			tv.mark_hits_per_lane(fast_slot, pr.y >> 24, wm, lane);
			if (fast_slot >= 0) verdict = TRI_SEPARATED;  // settled
			// pairs the filter could not separate: certain contacts skip the fp64 triangle test
			unsigned mm = __ballot_sync(FULL, verdict != TRI_SEPARATED);
			const unsigned sure = __ballot_sync(FULL, verdict == TRI_INTERSECT);
			while (mm) {
				const int from = __ffs(mm) - 1;
				mm &= mm - 1;
				const uint32_t t1 = __shfl_sync(FULL, pr.x, from), w2 = __shfl_sync(FULL, pr.y, from);
				const uint32_t t2 = w2 & 0xffffffu, cand = w2 >> 24;
				const unsigned src = __float_as_uint(wm.cand[rec_at(cand, 3)].y);  // pose slot of this candidate
				const unsigned e = tv.object(cand, wm);
				if (tv.dead(src, wm)) continue;
				if (COUNT && !(((sure >> from) & 1u) && !p.always_replay)) ctr[CTR_EXACT] += 1;
				// fp64 relative transform of this candidate, FCL's operation order: recomputed from the pose
				// record into shared memory (fused kernel) or read from the candidate list (split pipeline)
				const DevObject &ob = p.objects[e];
				const double *rel = tv.exact_rel(p, wm, ob, src, cand, lane);
				const DevModel &m1 = p.env.models[ob.model];
				const bool hit = ((sure >> from) & 1u) && !p.always_replay
				                     ? true
				                     : exact_tri_intersect_warp(rel, p.env.tri64 + 9ull * (m1.tri_base + t1),
				                                                p.agent.tri64 + 9ull * (m2.tri_base + t2), lane);
				if (hit) {
					// Path certificate.  FCL reports this contact only if every OBB pair on its descent to
					// the two leaves passed obbDisjoint.  Each of those boxes contains its triangle, the two
					// triangles were just found to meet, and obbDisjoint pads every radius by 1e-6 times (at
					// least) the two smallest extents of one box; when that pad (>= 1e-6 mu) exceeds the
					// magnitudes entering the test by a factor >= 1e5 above fp64 round-off (2^-53 ~ 1e-16),
					// no test on the path can report "disjoint" and the replay is provably redundant
					// (DESIGN.md "Path certificate").  mu = min over the two root-to-leaf chains, precomputed per
					// triangle; it is 0 for degenerate boxes, which forces the replay.
					const double mag = fabs(rel[9]) + fabs(rel[10]) + fabs(rel[11]) + m1.cmax1 + m2.cmax1 +
					                   m1.esum + m2.esum;
					const double mu = fmin((double)p.env.mu_chain[m1.tri_base + t1], (double)p.agent.mu_chain[m2.tri_base + t2]);
					const bool certified = !p.always_replay && mu >= 1e-5 * mag;
					if (certified) {
						tv.mark_hit(src, cand, wm, p, lane);
					} else {
						if (COUNT) ctr[CTR_PATH] += 1;
						if (exact_path_reaches(p, m1, m2, rel, t1, t2, lane)) tv.mark_hit(src, cand, wm, p, lane);
					}
				}
			}
			__syncwarp();
			continue;
		}
		// ---- 4. node step.  `cur` is a PLAN: a pair already known to overlap, stored as "the two children
		//         of the tree FCL descends next, against the other node".  One step fetches the sibling
		//         line (both children share 128 bytes) plus the other node and runs both tests: one memory
		//         round trip per two OBB tests.  Survivors become plans themselves; the first child's plan
		//         stays current and the second is stacked (FCL visits the first child first). -----------------
		bool leaf0 = false, leaf1 = false;
		uint2 lf0 = make_uint2(0u, 0u), lf1 = make_uint2(0u, 0u);
		unsigned n_tests = 0;
		if (has) {
			const uint2 pr = cur;
			has = false;
			const unsigned tag = pr.y & 0xff000000u;
			const unsigned nc = pr.y >> 24;
			const float4 r3 = wm.cand[rec_at(nc, 3)];
			if (tv.dead(__float_as_uint(r3.y), wm)) {
				top = bot;  // this lane's stack holds one candidate only, and its pose-set already collides
			} else {
				const float4 r0 = wm.cand[rec_at(nc, 0)], r1 = wm.cand[rec_at(nc, 1)], r2 = wm.cand[rec_at(nc, 2)];
				RelF x;
				x.R[0] = r0.x; x.R[1] = r0.y; x.R[2] = r0.z; x.R[3] = r0.w;
				x.R[4] = r1.x; x.R[5] = r1.y; x.R[6] = r1.z; x.R[7] = r1.w;
				x.R[8] = r2.x; x.T[0] = r2.y; x.T[1] = r2.z; x.T[2] = r2.w;
				x.S = r3.x;
				const float4 *nodes1 = p.env.node32 + 4ull * __float_as_uint(r3.z);
				const bool single = (pr.y & kPlanSingle) != 0u;           // root pair: one test
				const bool env_side = single || (pr.y & kPlanEnv) != 0u;  // the children are environment nodes
				const uint32_t i2base = pr.y & kPlanIdxMask;
				uint2 plan0 = make_uint2(0u, 0u), plan1 = make_uint2(0u, 0u);
				bool ok0 = false, ok1 = false;
				float d2_0 = 0.f, d2_1 = 0.f;
				// the node that stays (fetched once) and the sibling line of the tree that is descended
				const float4 *op = env_side ? nodes2 + 4ull * i2base : nodes1 + 4ull * pr.x;
				const float4 *cp = env_side ? nodes1 + 4ull * pr.x : nodes2 + 4ull * i2base;
				const F8 oA = ldg256(op), oB = ldg256(op + 2);
				const float4 o0 = oA.a, o1 = oA.b, o2 = oB.a, o3 = oB.b;
#pragma unroll
				for (int k = 0; k < 2; ++k) {
					if (k == 1 && single) break;
					const uint32_t i1 = pr.x + ((env_side && k) ? 1u : 0u);
					const uint32_t i2 = i2base + ((!env_side && k) ? 1u : 0u);
					const F8 cA = ldg256(cp + 4 * k), cB = ldg256(cp + 4 * k + 2);
					const float4 c0 = cA.a, c1 = cA.b, c2 = cB.a, c3 = cB.b;
					const float4 a0 = env_side ? c0 : o0, a1 = env_side ? c1 : o1, a2 = env_side ? c2 : o2,
					             a3 = env_side ? c3 : o3;
					const float4 b0 = env_side ? o0 : c0, b1 = env_side ? o1 : c1, bb2 = env_side ? o2 : c2,
					             b3 = env_side ? o3 : c3;
					++n_tests;
					float d2;
					if (obb_disjoint_f32(a0, a1, a2, a3, b0, b1, bb2, b3, x, d2)) continue;
					if (k == 0) d2_0 = d2; else d2_1 = d2;
					const int link1 = __float_as_int(a3.w), link2 = __float_as_int(b3.w);
					const bool l1 = link1 < 0, l2 = link2 < 0;
					if (l1 && l2) {
						const uint2 lf = make_uint2((uint32_t)(-(link1 + 1)), (uint32_t)(-(link2 + 1)) | tag);
						if (k == 0) { leaf0 = true; lf0 = lf; } else { leaf1 = true; lf1 = lf; }
						continue;
					}
					// firstOverSecond(i1, i2): descend the environment tree iff l2 || (!l1 && size1 > size2)
					const float s1 = a3.x * a3.x + a3.y * a3.y + a3.z * a3.z;
					const float s2 = b3.x * b3.x + b3.y * b3.y + b3.z * b3.z;
					const bool first = l2 || (!l1 && s1 > s2);
					const uint2 pl = first ? make_uint2((uint32_t)link1, i2 | kPlanEnv | tag)
					                       : make_uint2(i1, (uint32_t)link2 | tag);
					if (k == 0) { ok0 = true; plan0 = pl; } else { ok1 = true; plan1 = pl; }
				}
				if (ok0 && ok1) {
					// Both children overlap.  FCL would take child 0 first; the verdict does not depend on the
					// order, and descending first into the child whose box centre is closer to the other box
					// reaches a contact (if there is one) after fewer tests.
					const bool swap = p.near_first && d2_1 < d2_0;
					cur = swap ? plan1 : plan0;
					has = true;
					stk[top++ & M] = swap ? plan0 : plan1;
				} else if (ok0) {
					cur = plan0;
					has = true;
				} else if (ok1) {
					cur = plan1;
					has = true;
				}
			}
		}
		if (COUNT) {
			unsigned t = n_tests;
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
			ctr[CTR_BV] += t;
		}
		{
			const unsigned lm0 = __ballot_sync(FULL, leaf0);
			if (leaf0) wm.leafq[lq + __popc(lm0 & lt)] = lf0;
			lq += __popc(lm0);
		}
		const bool push_leaf = leaf1;
		const uint2 leaf = lf1;
		const unsigned lm = __ballot_sync(FULL, push_leaf);
		if (push_leaf) wm.leafq[lq + __popc(lm & lt)] = leaf;
		lq += __popc(lm);
		__syncwarp();
	}
}


}  // namespace
}  // namespace mptcc
