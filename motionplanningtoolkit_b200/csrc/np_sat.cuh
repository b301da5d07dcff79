// np_sat.cuh -- traversal node and the child-frame 15-axis SAT shared by the second- and third-generation
// narrowphase kernels (np2_kernels.cu, np3_kernels.cu).  Margins: DESIGN.md section 6.
#pragma once

#ifndef MPTCC_NP_SAT_HOST  // tests/host_sat.cpp compiles the arithmetic of this header for the CPU against its own stand-ins
#include <cuda_fp16.h>

#include "collide_device.cuh"
#endif

namespace mptcc {
namespace {

struct Node {
	float4 v0, v1, v2, v3;  // (axis0, c.x) (axis1, c.y) (axis2, c.z) (extent, link)
};
__device__ __forceinline__ Node node_global(const float4 *p) {
#ifdef MPTCC_NODE_EL
	const F8 a = ldg256_evict_last(p), b = ldg256_evict_last(p + 2);
#else
	const F8 a = ldg256(p), b = ldg256(p + 2);
#endif
	return Node{a.a, a.b, b.a, b.b};
}
__device__ __forceinline__ Node node_shared(const float4 *p) { return Node{p[0], p[1], p[2], p[3]}; }

// Quantised node (device_layout.hpp, nodeq) -> Node.  v3.w carries the link as the int the rest of the code expects.
// A 22-bit fixed-point component Q stands for q = Q * 2^-21 - 1.  The bits of Q, read as a float, are the subnormal
// Q * 2^-149, and fma(Q * 2^-149, 2^127, -0.5) = q / 2 exactly (sm_100 FFMA takes subnormal operands at full speed): a
// mask and one FFMA per component, no shift into a mantissa and no integer-to-float conversion.  The rotation matrix is
// quadratic in the quaternion, so the factor 1/2 per component turns its 2 into 8 -- every product below is the
// product of the unscaled formula times a power of two, i.e. bit-identical to it.
constexpr float kQHalfScale = 1.7014118346046923e38f;  // 2^127
__device__ __forceinline__ float qcomp(uint32_t Q) { return fmaf(__uint_as_float(Q), kQHalfScale, -0.5f); }
// the 24-bit two's-complement link: bytes 2, 3 of q1.x, byte 3 of q1.w, sign-extended by one PRMT
__device__ __forceinline__ int qlink(const uint4 q1) {
#ifdef __CUDA_ARCH__
	int r;
	asm("prmt.b32 %0, %1, %2, 0xF732;" : "=r"(r) : "r"(q1.x), "r"(q1.w));
	return r;
#else  // (the host compile of tests/host_sat.cpp)
	const uint32_t raw = (q1.x >> 16) | ((q1.w >> 24) << 16);
	return (int)(raw << 8) >> 8;
#endif
}
__device__ __forceinline__ Node node_decode(const uint4 q0, const uint4 q1) {
	constexpr uint32_t M22 = 0x3fffffu;
	const float w = qcomp(q1.y & M22), x = qcomp(__funnelshift_r(q1.y, q1.z, 22) & M22);
	const float y = qcomp(__funnelshift_r(q1.z, q1.w, 12) & M22), z = qcomp((q1.w >> 2) & M22);  // (halves of the components)
	const float2 e = __half22float2(*reinterpret_cast<const __half2 *>(&q0.w));  // extent.x, extent.y
	const float ez = __half2float(__ushort_as_half((unsigned short)(q1.x & 0xffffu)));
	const int link = qlink(q1);
	const float tx = 8.0f * x, ty = 8.0f * y, tz = 8.0f * z;
	const float twx = tx * w, twy = ty * w, twz = tz * w;
	const float txx = tx * x, txy = ty * x, txz = tz * x;
	const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
	Node n;
	n.v0 = make_float4(1.0f - (tyy + tzz), txy - twz, txz + twy, __uint_as_float(q0.x));
	n.v1 = make_float4(txy + twz, 1.0f - (txx + tzz), tyz - twx, __uint_as_float(q0.y));
	n.v2 = make_float4(txz - twy, tyz + twx, 1.0f - (txx + tyy), __uint_as_float(q0.z));
	n.v3 = make_float4(e.x, e.y, ez, __int_as_float(link));
	return n;
}
__device__ __forceinline__ Node qnode_global(const uint4 *p) {
	uint4 q0, q1;
	asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	    : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w), "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w)
	    : "l"(p));
	return node_decode(q0, q1);
}
__device__ __forceinline__ Node qnode_shared(const uint4 *p) { return node_decode(p[0], p[1]); }
// A sibling pair (64 contiguous bytes) with two 256-bit loads issued back to back.

// The 15-axis SAT in the frame of box a (extents ea) against box b (extents eb): G(i,j) = a_i . b_j, t = centre of
// b in a's frame.  One-sided: true only if the boxes are certainly disjoint (margins: DESIGN.md section 6).
__device__ __forceinline__ bool sat15(const float *G, const float t0, const float t1, const float t2, const float4 ea,
                                      const float4 eb, const float mg) {
	float F[9];
#pragma unroll
	for (int k = 0; k < 9; ++k) F[k] = fabsf(G[k]) + 3e-6f;
	bool dj = false;
	dj |= fabsf(t0) > ea.x + (F[0] * eb.x + F[1] * eb.y + F[2] * eb.z) + mg;
	dj |= fabsf(t1) > ea.y + (F[3] * eb.x + F[4] * eb.y + F[5] * eb.z) + mg;
	dj |= fabsf(t2) > ea.z + (F[6] * eb.x + F[7] * eb.y + F[8] * eb.z) + mg;
	dj |= fabsf(t0 * G[0] + t1 * G[3] + t2 * G[6]) > eb.x + (F[0] * ea.x + F[3] * ea.y + F[6] * ea.z) + mg;
	dj |= fabsf(t0 * G[1] + t1 * G[4] + t2 * G[7]) > eb.y + (F[1] * ea.x + F[4] * ea.y + F[7] * ea.z) + mg;
	dj |= fabsf(t0 * G[2] + t1 * G[5] + t2 * G[8]) > eb.z + (F[2] * ea.x + F[5] * ea.y + F[8] * ea.z) + mg;
	// a0 x b0..2
	dj |= fabsf(t2 * G[3] - t1 * G[6]) > ea.y * F[6] + ea.z * F[3] + eb.y * F[2] + eb.z * F[1] + mg;
	dj |= fabsf(t2 * G[4] - t1 * G[7]) > ea.y * F[7] + ea.z * F[4] + eb.x * F[2] + eb.z * F[0] + mg;
	dj |= fabsf(t2 * G[5] - t1 * G[8]) > ea.y * F[8] + ea.z * F[5] + eb.x * F[1] + eb.y * F[0] + mg;
	// a1 x b0..2
	dj |= fabsf(t0 * G[6] - t2 * G[0]) > ea.x * F[6] + ea.z * F[0] + eb.y * F[5] + eb.z * F[4] + mg;
	dj |= fabsf(t0 * G[7] - t2 * G[1]) > ea.x * F[7] + ea.z * F[1] + eb.x * F[5] + eb.z * F[3] + mg;
	dj |= fabsf(t0 * G[8] - t2 * G[2]) > ea.x * F[8] + ea.z * F[2] + eb.x * F[4] + eb.y * F[3] + mg;
	// a2 x b0..2
	dj |= fabsf(t1 * G[0] - t0 * G[3]) > ea.x * F[3] + ea.y * F[0] + eb.y * F[8] + eb.z * F[7] + mg;
	dj |= fabsf(t1 * G[1] - t0 * G[4]) > ea.x * F[4] + ea.y * F[1] + eb.x * F[8] + eb.z * F[6] + mg;
	dj |= fabsf(t1 * G[2] - t0 * G[5]) > ea.x * F[5] + ea.y * F[2] + eb.x * F[7] + eb.y * F[6] + mg;
	return dj;
}

// Stationary half of a step: W = Ry [S axes as columns] (W[3r+j] = Ry.row(r) . s_j), w = Ry cS + Ty.
struct StepFrame {
	float W[9], w[3];
	float s1;  // |Ty|_1 + |cS|_1
};

// The frame of a step from the candidate record r0..r3 = R (9), Tc (3), Tinv (3), S: (Ry, Ty) is the relative transform
// (R, Tc) when the children are environment nodes (env_side), its inverse (R^T, Tinv) when they are agent nodes.
__device__ __forceinline__ StepFrame step_frame(const float4 r0, const float4 r1, const float4 r2, const float4 r3, const bool env_side,
                                                const Node &S) {
	StepFrame f;
	const float ti0 = r3.x, ti1 = r3.y, ti2 = r3.z;
	const float y00 = r0.x, y11 = r1.x, y22 = r2.x;
	const float y01 = env_side ? r0.y : r0.w, y10 = env_side ? r0.w : r0.y;
	const float y02 = env_side ? r0.z : r1.z, y20 = env_side ? r1.z : r0.z;
	const float y12 = env_side ? r1.y : r1.w, y21 = env_side ? r1.w : r1.y;
	const float ty0 = env_side ? r2.y : ti0, ty1 = env_side ? r2.z : ti1, ty2 = env_side ? r2.w : ti2;
	f.W[0] = y00 * S.v0.x + y01 * S.v0.y + y02 * S.v0.z;
	f.W[1] = y00 * S.v1.x + y01 * S.v1.y + y02 * S.v1.z;
	f.W[2] = y00 * S.v2.x + y01 * S.v2.y + y02 * S.v2.z;
	f.W[3] = y10 * S.v0.x + y11 * S.v0.y + y12 * S.v0.z;
	f.W[4] = y10 * S.v1.x + y11 * S.v1.y + y12 * S.v1.z;
	f.W[5] = y10 * S.v2.x + y11 * S.v2.y + y12 * S.v2.z;
	f.W[6] = y20 * S.v0.x + y21 * S.v0.y + y22 * S.v0.z;
	f.W[7] = y20 * S.v1.x + y21 * S.v1.y + y22 * S.v1.z;
	f.W[8] = y20 * S.v2.x + y21 * S.v2.y + y22 * S.v2.z;
	const float cx = S.v0.w, cy = S.v1.w, cz = S.v2.w;
	f.w[0] = y00 * cx + y01 * cy + y02 * cz + ty0;
	f.w[1] = y10 * cx + y11 * cy + y12 * cz + ty1;
	f.w[2] = y20 * cx + y21 * cy + y22 * cz + ty2;
	f.s1 = (fabsf(ty0) + fabsf(ty1) + fabsf(ty2)) + (fabsf(cx) + fabsf(cy) + fabsf(cz));
	return f;
}

// Q: the decoded axes of a quantised node are within 4e-6 per component of the true ones; the stored extents are
// widened so that the box they span contains FCL's box, and this term covers what the SAT's algebra loses because the
// decoded frames are orthonormal only to ~1e-5.
constexpr float kQFrameMargin = 6.4e-5f;

// Child half: true => child box C and stationary box S are certainly disjoint.  d2 = squared centre distance.
template <bool Q>
__device__ __forceinline__ bool child_disjoint(const StepFrame &f, const Node &C, const float4 es, float &d2) {
	float G[9];
#pragma unroll
	for (int j = 0; j < 3; ++j) {
		G[0 + j] = C.v0.x * f.W[j] + C.v0.y * f.W[3 + j] + C.v0.z * f.W[6 + j];
		G[3 + j] = C.v1.x * f.W[j] + C.v1.y * f.W[3 + j] + C.v1.z * f.W[6 + j];
		G[6 + j] = C.v2.x * f.W[j] + C.v2.y * f.W[3 + j] + C.v2.z * f.W[6 + j];
	}
	const float dx = f.w[0] - C.v0.w, dy = f.w[1] - C.v1.w, dz = f.w[2] - C.v2.w;
	d2 = dx * dx + dy * dy + dz * dz;
	const float t0 = C.v0.x * dx + C.v0.y * dy + C.v0.z * dz;
	const float t1 = C.v1.x * dx + C.v1.y * dy + C.v1.z * dz;
	const float t2 = C.v2.x * dx + C.v2.y * dy + C.v2.z * dz;
	// margin = 192 u (|Ty|_1 + |cS|_1 + |cC|_1)
	float mg = (192.0f * kU) * (f.s1 + (fabsf(C.v0.w) + fabsf(C.v1.w) + fabsf(C.v2.w)));
	if (Q) mg += kQFrameMargin * ((C.v3.x + C.v3.y + C.v3.z) + (es.x + es.y + es.z));
	return sat15(G, t0, t1, t2, C.v3, es, mg);
}

// ------------------------------------------------------------------------------------------------------------------
// Two-wide evaluation: a traversal step tests the two children of one node against the same stationary box, so the
// two SATs are the same arithmetic on two data sets.  sm_100 has two-wide fp32 instructions (FFMA2 / FMUL2 / FADD2 on
// 64-bit register pairs, scalar operands broadcast for free); child 0 lives in the .x halves and child 1 in the .y
// halves.  Same formulas and margins as child_disjoint / sat15 above.
// ------------------------------------------------------------------------------------------------------------------
struct Node2 {
	float2 a[9];  // a[3 i + k] = component k of axis i
	float2 c[3];  // centre
	float2 e[3];  // extents
	int link0, link1;
};
__device__ __forceinline__ float2 bc2(float s) { return make_float2(s, s); }
__device__ __forceinline__ float2 abs2(float2 v) { return make_float2(fabsf(v.x), fabsf(v.y)); }
__device__ __forceinline__ float2 neg2(float2 v) { return make_float2(-v.x, -v.y); }
#define MPTCC_MUL2 __fmul2_rn
#define MPTCC_FMA2 __ffma2_rn
#define MPTCC_ADD2 __fadd2_rn

__device__ __forceinline__ Node2 pack_nodes(const Node &A, const Node &B) {
	Node2 n;
	n.a[0] = make_float2(A.v0.x, B.v0.x); n.a[1] = make_float2(A.v0.y, B.v0.y); n.a[2] = make_float2(A.v0.z, B.v0.z);
	n.a[3] = make_float2(A.v1.x, B.v1.x); n.a[4] = make_float2(A.v1.y, B.v1.y); n.a[5] = make_float2(A.v1.z, B.v1.z);
	n.a[6] = make_float2(A.v2.x, B.v2.x); n.a[7] = make_float2(A.v2.y, B.v2.y); n.a[8] = make_float2(A.v2.z, B.v2.z);
	n.c[0] = make_float2(A.v0.w, B.v0.w); n.c[1] = make_float2(A.v1.w, B.v1.w); n.c[2] = make_float2(A.v2.w, B.v2.w);
	n.e[0] = make_float2(A.v3.x, B.v3.x); n.e[1] = make_float2(A.v3.y, B.v3.y); n.e[2] = make_float2(A.v3.z, B.v3.z);
	n.link0 = __float_as_int(A.v3.w);
	n.link1 = __float_as_int(B.v3.w);
	return n;
}

// Two quantised nodes (device_layout.hpp, nodeq) decoded side by side.
__device__ __forceinline__ float2 qcomp2(uint32_t Qa, uint32_t Qb) {
	return MPTCC_FMA2(make_float2(__uint_as_float(Qa), __uint_as_float(Qb)), bc2(kQHalfScale), bc2(-0.5f));
}
__device__ __forceinline__ Node2 node2_decode(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1) {
	constexpr uint32_t M22 = 0x3fffffu;
	const float2 w = qcomp2(a1.y & M22, b1.y & M22);
	const float2 x = qcomp2(__funnelshift_r(a1.y, a1.z, 22) & M22, __funnelshift_r(b1.y, b1.z, 22) & M22);
	const float2 y = qcomp2(__funnelshift_r(a1.z, a1.w, 12) & M22, __funnelshift_r(b1.z, b1.w, 12) & M22);
	const float2 z = qcomp2((a1.w >> 2) & M22, (b1.w >> 2) & M22);
	const float2 tx = MPTCC_MUL2(x, bc2(8.0f)), ty = MPTCC_MUL2(y, bc2(8.0f)), tz = MPTCC_MUL2(z, bc2(8.0f));
	const float2 twx = MPTCC_MUL2(tx, w), twy = MPTCC_MUL2(ty, w), twz = MPTCC_MUL2(tz, w);
	const float2 txx = MPTCC_MUL2(tx, x), tyy = MPTCC_MUL2(ty, y), tzz = MPTCC_MUL2(tz, z);
	const float2 one = bc2(1.0f), m1 = bc2(-1.0f);
	Node2 n;
	n.a[0] = MPTCC_FMA2(MPTCC_ADD2(tyy, tzz), m1, one);
	n.a[1] = MPTCC_FMA2(ty, x, neg2(twz));
	n.a[2] = MPTCC_FMA2(tz, x, twy);
	n.a[3] = MPTCC_FMA2(ty, x, twz);
	n.a[4] = MPTCC_FMA2(MPTCC_ADD2(txx, tzz), m1, one);
	n.a[5] = MPTCC_FMA2(tz, y, neg2(twx));
	n.a[6] = MPTCC_FMA2(tz, x, neg2(twy));
	n.a[7] = MPTCC_FMA2(tz, y, twx);
	n.a[8] = MPTCC_FMA2(MPTCC_ADD2(txx, tyy), m1, one);
	n.c[0] = make_float2(__uint_as_float(a0.x), __uint_as_float(b0.x));
	n.c[1] = make_float2(__uint_as_float(a0.y), __uint_as_float(b0.y));
	n.c[2] = make_float2(__uint_as_float(a0.z), __uint_as_float(b0.z));
	const float2 ea = __half22float2(*reinterpret_cast<const __half2 *>(&a0.w)), eb = __half22float2(*reinterpret_cast<const __half2 *>(&b0.w));
	n.e[0] = make_float2(ea.x, eb.x);
	n.e[1] = make_float2(ea.y, eb.y);
	n.e[2] = make_float2(__half2float(__ushort_as_half((unsigned short)(a1.x & 0xffffu))), __half2float(__ushort_as_half((unsigned short)(b1.x & 0xffffu))));
	n.link0 = qlink(a1);
	n.link1 = qlink(b1);
	return n;
}
__device__ __forceinline__ void ldg256u(const uint4 *p, uint4 &q0, uint4 &q1) {
	asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	    : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w), "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w)
	    : "l"(p));
}

// Both children against the stationary box: bit 0 / bit 1 set = child 0 / child 1 certainly disjoint from S.
template <bool Q>
__device__ __forceinline__ unsigned pair_disjoint(const StepFrame &f, const Node2 &C, const float4 es, float2 &d2) {
	float2 G[9];
#pragma unroll
	for (int i = 0; i < 3; ++i)
#pragma unroll
		for (int j = 0; j < 3; ++j)
			G[3 * i + j] = MPTCC_FMA2(C.a[3 * i + 2], bc2(f.W[6 + j]), MPTCC_FMA2(C.a[3 * i + 1], bc2(f.W[3 + j]), MPTCC_MUL2(C.a[3 * i], bc2(f.W[j]))));
	const float2 m1 = bc2(-1.0f);
	const float2 dx = MPTCC_FMA2(C.c[0], m1, bc2(f.w[0])), dy = MPTCC_FMA2(C.c[1], m1, bc2(f.w[1])), dz = MPTCC_FMA2(C.c[2], m1, bc2(f.w[2]));
	d2 = MPTCC_FMA2(dz, dz, MPTCC_FMA2(dy, dy, MPTCC_MUL2(dx, dx)));
	float2 t[3], nt[3];
#pragma unroll
	for (int i = 0; i < 3; ++i) {
		t[i] = MPTCC_FMA2(C.a[3 * i + 2], dz, MPTCC_FMA2(C.a[3 * i + 1], dy, MPTCC_MUL2(C.a[3 * i], dx)));
		nt[i] = neg2(t[i]);
	}
	// margin = 192 u (|Ty|_1 + |cS|_1 + |cC|_1)
	float2 mg = MPTCC_MUL2(bc2(192.0f * kU), MPTCC_ADD2(bc2(f.s1), MPTCC_ADD2(MPTCC_ADD2(abs2(C.c[0]), abs2(C.c[1])), abs2(C.c[2]))));
	// |G| + 3e-6 of sat15 (the pad over-covers the rounding of G): the pad is pulled out of the sums -- every radius sum
	// multiplies it with a subset of the six extents, so 3e-6 (sum of all six) in the margin is at least as large -- and
	// |G| is an operand modifier of FFMA2, which saves nine additions and nine register pairs.
	const float2 esum = MPTCC_ADD2(MPTCC_ADD2(MPTCC_ADD2(C.e[0], C.e[1]), C.e[2]), bc2(es.x + es.y + es.z));
	if (Q) mg = MPTCC_FMA2(bc2(kQFrameMargin + 3e-6f), esum, mg);
	else mg = MPTCC_FMA2(bc2(3e-6f), esum, mg);
	float2 F[9];
#pragma unroll
	for (int k = 0; k < 9; ++k) F[k] = abs2(G[k]);
	const float ebv[3] = {es.x, es.y, es.z};
	unsigned dj = 0u;
#define MPTCC_CMP2(lhs, rhs)                                                       \
	{                                                                              \
		const float2 l_ = (lhs), r_ = (rhs);                                       \
		dj |= (fabsf(l_.x) > r_.x ? 1u : 0u) | (fabsf(l_.y) > r_.y ? 2u : 0u);     \
	}
	// face axes of the child: |t_i| > ea_i + (F[3i] eb0 + F[3i+1] eb1 + F[3i+2] eb2) + mg
#pragma unroll
	for (int i = 0; i < 3; ++i)
		MPTCC_CMP2(t[i], MPTCC_FMA2(F[3 * i + 2], bc2(ebv[2]), MPTCC_FMA2(F[3 * i + 1], bc2(ebv[1]), MPTCC_FMA2(F[3 * i], bc2(ebv[0]), MPTCC_ADD2(C.e[i], mg)))))
	// face axes of S: |t . G(:,j)| > eb_j + (F[j] ea0 + F[3+j] ea1 + F[6+j] ea2) + mg
#pragma unroll
	for (int j = 0; j < 3; ++j)
		MPTCC_CMP2(MPTCC_FMA2(t[2], G[6 + j], MPTCC_FMA2(t[1], G[3 + j], MPTCC_MUL2(t[0], G[j]))),
		           MPTCC_FMA2(F[6 + j], C.e[2], MPTCC_FMA2(F[3 + j], C.e[1], MPTCC_FMA2(F[j], C.e[0], MPTCC_ADD2(bc2(ebv[j]), mg)))))
	// edge axes a_i x b_j, (i1, i2) = the other two child axes, (j1, j2) = the other two axes of S:
	//   |t_i2 G(i1,j) - t_i1 G(i2,j)| > ea_i1 F(i2,j) + ea_i2 F(i1,j) + eb_j1 F(i,j2) + eb_j2 F(i,j1) + mg
#ifndef MPTCC_SAT_EDGE_MASK
// bit 3 i + j: test axis a_i x b_j.  A subset only makes the filter more conservative (verdicts do not depend on it).
// Measured on C5 (tools/ab_np.py, gpurun_out/ab19-20): all nine 1.040 ms at 51.2 tests per pose-set, none 1.033 ms at
// 58.6; the four axes between the two longest axes of each box (FCL's fit puts the direction of largest spread first, the thinnest last) 0.997 ms at 52.2.
#define MPTCC_SAT_EDGE_MASK 0x1b
#endif
#pragma unroll
	for (int i = 0; i < 3; ++i) {
		const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
		for (int j = 0; j < 3; ++j) {
			const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
			if ((MPTCC_SAT_EDGE_MASK >> (3 * i + j)) & 1)
				MPTCC_CMP2(MPTCC_FMA2(t[i2], G[3 * i1 + j], MPTCC_MUL2(nt[i1], G[3 * i2 + j])),
				           MPTCC_FMA2(F[3 * i + j1], bc2(ebv[j2]), MPTCC_FMA2(F[3 * i + j2], bc2(ebv[j1]), MPTCC_FMA2(C.e[i2], F[3 * i1 + j], MPTCC_FMA2(C.e[i1], F[3 * i2 + j], mg)))))
		}
	}
#undef MPTCC_CMP2
	return dj;
}

}  // namespace
}  // namespace mptcc
