// host_sat.cpp -- TEST INFRASTRUCTURE: the arithmetic of motionplanningtoolkit_b200/csrc/np_sat.cuh (quantised-node
// decode, step frame, two-wide 11-axis SAT with its margins) compiled for the CPU, so that the one-sidedness of the
// narrowphase's fp32 filter ("disjoint" only if FCL's fp64 obbDisjoint says so too) can be checked against the oracle
// without a GPU (tests/test_host_library.py).  Only stand-ins for the CUDA vector types and intrinsics live here; the
// formulas are the header's own.  The device contracts a*b+c into FMAs where this build does not (and vice versa):
// both evaluations are inside the error analysis of DESIGN.md section 6, which is what the test exercises.
//   g++ -O1 -ffp-contract=off -shared -fPIC -I motionplanningtoolkit_b200/csrc tests/host_sat.cpp -o host_sat.so
#include <cmath>
#include <cstdint>
#include <cstring>

#define MPTCC_NP_SAT_HOST 1
#define __device__
#define __forceinline__ inline

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (s & 31u)); }
// two-wide fp32: each half rounded once, like FMUL2 / FFMA2 / FADD2
static inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
// binary16 -> float, exact
struct __half { uint16_t v; };
struct __half2 { uint16_t lo, hi; };
static inline float half_bits_to_float(uint16_t h) {
	const uint32_t s = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 31u, m = h & 1023u;
	if (e == 0) return (s ? -1.0f : 1.0f) * std::ldexp((float)m, -24);
	if (e == 31) return __uint_as_float(s | 0x7f800000u | (m << 13));
	return __uint_as_float(s | ((e + 112u) << 23) | (m << 13));
}
static inline __half __ushort_as_half(unsigned short v) { return __half{v}; }
static inline float __half2float(__half h) { return half_bits_to_float(h.v); }
static inline float2 __half22float2(__half2 h) { return float2{half_bits_to_float(h.lo), half_bits_to_float(h.hi)}; }

namespace mptcc {
namespace {
constexpr float kU = 5.9604645e-8f;  // 2^-24 (collide_device.cuh)
struct F8 { float4 a, b; };
inline F8 ldg256(const float4 *p) { return F8{p[0], p[1]}; }
}  // namespace
}  // namespace mptcc

#include "np_sat.cuh"

extern "C" {
// s, c0, c1: the eight words of three quantised nodes (stationary node, the two children); rec16 = the candidate
// record r0..r3; returns bit 0 / bit 1 = child 0 / child 1 certainly disjoint from the stationary node.
unsigned host_pair_disjoint_q(const uint32_t *s, const uint32_t *c0, const uint32_t *c1, const float *rec16, int env_side) {
	using namespace mptcc;
	const Node S = node_decode(uint4{s[0], s[1], s[2], s[3]}, uint4{s[4], s[5], s[6], s[7]});
	const Node2 C2 = node2_decode(uint4{c0[0], c0[1], c0[2], c0[3]}, uint4{c0[4], c0[5], c0[6], c0[7]}, uint4{c1[0], c1[1], c1[2], c1[3]},
	                              uint4{c1[4], c1[5], c1[6], c1[7]});
	const float4 r0{rec16[0], rec16[1], rec16[2], rec16[3]}, r1{rec16[4], rec16[5], rec16[6], rec16[7]};
	const float4 r2{rec16[8], rec16[9], rec16[10], rec16[11]}, r3{rec16[12], rec16[13], rec16[14], rec16[15]};
	const StepFrame f = step_frame(r0, r1, r2, r3, env_side != 0, S);
	float2 d2;
	return pair_disjoint<true>(f, C2, S.v3, d2);
}
// the same step on 64-byte fp32 nodes (np3 / np2 with qnode = 0): s, c0, c1 = four float4 each (node32 layout)
static mptcc::Node node_of(const float *p) {
	const float4 *q = reinterpret_cast<const float4 *>(p);
	return mptcc::Node{q[0], q[1], q[2], q[3]};
}
unsigned host_pair_disjoint_f32(const float *s, const float *c0, const float *c1, const float *rec16, int env_side) {
	using namespace mptcc;
	const Node S = node_of(s);
	const Node2 C2 = pack_nodes(node_of(c0), node_of(c1));
	const float4 r0{rec16[0], rec16[1], rec16[2], rec16[3]}, r1{rec16[4], rec16[5], rec16[6], rec16[7]};
	const float4 r2{rec16[8], rec16[9], rec16[10], rec16[11]}, r3{rec16[12], rec16[13], rec16[14], rec16[15]};
	const StepFrame f = step_frame(r0, r1, r2, r3, env_side != 0, S);
	float2 d2;
	return pair_disjoint<false>(f, C2, S.v3, d2);
}
// the scalar 15-axis test of the np2 kernel (child_disjoint / sat15) on the two children, fp32 nodes or quantised nodes
unsigned host_child_disjoint(const void *s, const void *c0, const void *c1, const float *rec16, int env_side, int quantised) {
	using namespace mptcc;
	Node S, C[2];
	if (quantised) {
		const uint32_t *w[3] = {static_cast<const uint32_t *>(s), static_cast<const uint32_t *>(c0), static_cast<const uint32_t *>(c1)};
		Node *dst[3] = {&S, &C[0], &C[1]};
		for (int k = 0; k < 3; ++k) *dst[k] = node_decode(uint4{w[k][0], w[k][1], w[k][2], w[k][3]}, uint4{w[k][4], w[k][5], w[k][6], w[k][7]});
	} else {
		S = node_of(static_cast<const float *>(s));
		C[0] = node_of(static_cast<const float *>(c0));
		C[1] = node_of(static_cast<const float *>(c1));
	}
	const float4 r0{rec16[0], rec16[1], rec16[2], rec16[3]}, r1{rec16[4], rec16[5], rec16[6], rec16[7]};
	const float4 r2{rec16[8], rec16[9], rec16[10], rec16[11]}, r3{rec16[12], rec16[13], rec16[14], rec16[15]};
	const StepFrame f = step_frame(r0, r1, r2, r3, env_side != 0, S);
	float d2;
	unsigned m = 0u;
	for (int k = 0; k < 2; ++k)
		m |= (quantised ? child_disjoint<true>(f, C[k], S.v3, d2) : child_disjoint<false>(f, C[k], S.v3, d2)) ? 1u << k : 0u;
	return m;
}
// the scalar decode and the two-wide decode must agree: returns the largest difference over axes, centres and extents
float host_decode_difference(const uint32_t *a, const uint32_t *b) {
	using namespace mptcc;
	const uint4 a0{a[0], a[1], a[2], a[3]}, a1{a[4], a[5], a[6], a[7]}, b0{b[0], b[1], b[2], b[3]}, b1{b[4], b[5], b[6], b[7]};
	const Node A = node_decode(a0, a1), B = node_decode(b0, b1);
	const Node2 P = node2_decode(a0, a1, b0, b1), Q = pack_nodes(A, B);
	float d = 0.f;
	for (int k = 0; k < 9; ++k) d = std::fmax(d, std::fmax(std::fabs(P.a[k].x - Q.a[k].x), std::fabs(P.a[k].y - Q.a[k].y)));
	for (int k = 0; k < 3; ++k) {
		d = std::fmax(d, std::fmax(std::fabs(P.c[k].x - Q.c[k].x), std::fabs(P.c[k].y - Q.c[k].y)));
		d = std::fmax(d, std::fmax(std::fabs(P.e[k].x - Q.e[k].x), std::fabs(P.e[k].y - Q.e[k].y)));
	}
	return (P.link0 == Q.link0 && P.link1 == Q.link1) ? d : 1e30f;
}
}
