// np_sat.cuh -- traversal node and the child-frame 15-axis SAT shared by the second- and third-generation
// narrowphase kernels (np2_kernels.cu, np3_kernels.cu).  Margins: DESIGN.md section 6.
#pragma once

#include <cuda_fp16.h>

#include "collide_device.cuh"

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
__device__ __forceinline__ Node node_decode(const uint4 q0, const uint4 q1) {
	const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&q0.x));  // axis0.x, axis0.y
	const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&q0.y));  // axis0.z, axis1.x
	const float2 c = __half22float2(*reinterpret_cast<const __half2 *>(&q0.z));  // axis1.y, axis1.z
	const float2 e = __half22float2(*reinterpret_cast<const __half2 *>(&q0.w));  // extent.x, extent.y
	const float ez = __half2float(__ushort_as_half((unsigned short)(q1.x & 0xffffu)));
	const uint32_t raw = (q1.x >> 16) | ((q1.y & 0xffu) << 16);
	const int link = (raw & 0x800000u) ? -(int)(raw & 0x7fffffu) - 1 : (int)raw;
	Node n;
	n.v0 = make_float4(a.x, a.y, b.x, __uint_as_float(q1.y));
	n.v1 = make_float4(b.y, c.x, c.y, __uint_as_float(q1.z));
	n.v2 = make_float4(a.y * c.y - b.x * c.x, b.x * b.y - a.x * c.y, a.x * c.x - a.y * b.y, __uint_as_float(q1.w));
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

// Child half: true => child box C and stationary box S are certainly disjoint.  d2 = squared centre distance.
// Q: the nodes come from the quantised array.  Their axes are off by <= 2^-12 per component (the third, a cross
// product, by twice that), so every projection radius the SAT forms from them can be short by a few 1e-4 of the
// box's size; 8e-3 (esum_C + esum_S) added to the margin covers that with a factor > 2.
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
	if (Q) mg += 8e-3f * ((C.v3.x + C.v3.y + C.v3.z) + (es.x + es.y + es.z));
	return sat15(G, t0, t1, t2, C.v3, es, mg);
}

}  // namespace
}  // namespace mptcc
