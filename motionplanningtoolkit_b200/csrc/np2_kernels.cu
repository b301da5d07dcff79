// np2_kernels.cu -- narrowphase of the split pipeline, second generation (round 2), plus the exact-pass kernel.
//
//   narrowphase2_kernel : persistent warps take batches of candidates (pose-set, environment object) from the list
//                         the broadphase kernel wrote and run the conservative fp32 BVH-vs-BVH traversal.
//   exact_kernel        : re-decides, in fp64 with FCL's operation order, the few leaf pairs the fp32 triangle
//                         filter could not settle (and replays FCL's OBB path where the certificate does not hold).
//
// What changed against narrowphase_kernel (split_kernels.cu), and why:
//   * TMA-staged tree top.  The agent's nodes are numbered breadth-first (HostPack::add, bfs = true), so the top
//     levels are one contiguous block; one elected thread per CTA pulls them into shared memory with
//     cp.async.bulk (SASS UBLKCP) completing on an mbarrier.  Agent-side node fetches with index < agent_top are
//     LDS instead of per-lane global gathers (the L1 data pipe, one wavefront per 32-byte sector per lane, is
//     what the round-1 kernel ran out of).
//   * Child-frame SAT with a shared half.  A traversal step tests the two children of one tree against one node
//     of the other.  The test is evaluated in the frame of the CHILD box: W = Ry [S axes], w = Ry cS + Ty are
//     computed once per step from the stationary node S, then per child G = C^T W, t = C^T (w - cC).  When the
//     children are agent nodes, (Ry, Ty) is the inverse relative transform (R^T, -R^T T), stored in the candidate
//     record by the broadphase kernel (fp64, rounded once).  One code path for both descent directions: no
//     role selects (46 of 302 instructions per test in round 1), 36 instead of 75 set-up FMAs per test.
//     The inverse is only the inverse for a rigid R; the broadphase kernel measures |R^T R - I| in fp64 and hands
//     launches with a non-rigid pose to the fused kernel (overflow flag = 2).
//   * No fp64 in the hot kernel.  Leaf pairs that need FCL's exact arithmetic are appended to a queue and
//     processed by exact_kernel afterwards, so the traversal kernel carries no fp64 code, no fp64 shared
//     scratch and fewer registers.
#include "collide_device.cuh"
#include "np_sat.cuh"
#include "launch_cache.hpp"

namespace mptcc {
namespace {

#ifndef MPTCC_NP2_WARPS
#define MPTCC_NP2_WARPS 16
#endif
#ifndef MPTCC_NP2_CTAS
#define MPTCC_NP2_CTAS 1
#endif
#ifndef MPTCC_NP2_BATCH
#define MPTCC_NP2_BATCH 96
#endif
#ifndef MPTCC_NP2_BATCH_LONG
#define MPTCC_NP2_BATCH_LONG 128
#endif
constexpr int kNp2Warps = MPTCC_NP2_WARPS;
constexpr int kNp2Batch = MPTCC_NP2_BATCH, kNp2BatchLong = MPTCC_NP2_BATCH_LONG;
constexpr unsigned kTopBytes = kAgentTopMaxQ * 32u > kAgentTopMax * 64u ? kAgentTopMaxQ * 32u : kAgentTopMax * 64u;  // staged agent nodes
constexpr unsigned kTopRegion = (kTopBytes + 8u + 127u) & ~127u;  // + the mbarrier, rounded for alignment

struct Np2Mem {
	uint2 *leafq;             // [kLeafQueueCap] (env prim, agent prim | cand << 24)
	float4 *cand;             // [BATCH][4] R(9) T(3) Tinv(3) S, swizzled with rec_at
	uint2 *cand_base;         // [BATCH] node_base, tri_base of the candidate's environment model
	uint32_t *cand_pose;      // [BATCH] pose-set index
	unsigned char *cand_slot; // [BATCH] slot = batch index of the first candidate of the same pose-set
	unsigned char *dead;      // [BATCH] per slot: pose-set already in collision
	unsigned char *donor;     // [32]
};

constexpr unsigned kLeafQueueCapX4 = 160;  // 32 waiting + up to 4 x 32 appended by one step
__host__ __device__ constexpr unsigned np2_warp_smem_bytes(unsigned batch, bool x4) {
	return (x4 ? kLeafQueueCapX4 : (unsigned)kLeafQueueCap) * 8u + batch * 64u + batch * 8u + batch * 4u + batch + batch + 32u;
}

// X4: a step expands BOTH trees at once -- the two children of the environment node against the two children of the
// agent node, four tests from four node fetches -- instead of FCL's one-tree-at-a-time descent (two tests from three
// fetches).  Half as many dependent memory round trips per candidate, a third fewer node bytes per test, and the
// per-step bookkeeping (record fetch, plan logic, work acquisition) is spread over four tests.  All tests are
// evaluated in the environment child's frame, so no inverse transform is needed.  Any traversal that only prunes
// certainly-disjoint pairs visits a superset of FCL's leaf pairs, so the verdicts do not depend on the expansion.
template <bool COUNT, int BATCH, bool Q, bool X4>
__device__ void traverse2(const KernelParams &p, const Np2Mem &wm, const float4 *__restrict__ atop, const DevModel &m2,
                          int ncand, unsigned batch_first, int lane, unsigned long long *ctr) {
	const float4 *__restrict__ nodes2 = p.agent.node32 + 4ull * m2.node_base;
	const uint4 *__restrict__ qnodes2 = p.agent.nodeq + 2ull * m2.node_base;
	const uint4 *__restrict__ qtop = reinterpret_cast<const uint4 *>(atop);
	const float4 *__restrict__ tris2 = p.agent.tri32 + 3ull * m2.tri_base;
	const unsigned lt = (1u << lane) - 1u;
	const unsigned ktop = p.agent_top;
	constexpr unsigned kStack = X4 ? 2u * kLaneStackCap : (unsigned)kLaneStackCap;  // (3 pending pairs per level with X4)
	constexpr unsigned M = kStack - 1;

	uint2 stk[kStack];
	unsigned top = 0, bot = 0;
	uint2 cur = make_uint2(0u, 0u);
	bool has = false;
	int qhead = 0, lq = 0;

	while (true) {
		if (!has && top != bot) {
			cur = stk[--top & M];
			has = true;
		}
		// ---- 1. idle lanes start the next unstarted candidates ------------------------------------------------
		unsigned im = __ballot_sync(FULL, !has);
		if (im && qhead < ncand) {
			const int c = qhead + __popc(im & lt);
			if (!has && c < ncand) {
				cur = make_uint2(0u, (X4 ? 0u : kPlanSingle) | ((unsigned)c << 24));
				has = true;
			}
			qhead += __popc(im);
			im = __ballot_sync(FULL, !has);
		}
		// ---- 2. ... or steal the top pending plan of a busy lane ------------------------------------------------
		if (im && qhead >= ncand && !p.no_steal) {
			const unsigned dm = __ballot_sync(FULL, has && top != bot);
			if (dm) {
				const int n_thieves = __popc(im), n_donors = __popc(dm);
				const int drank = __popc(dm & lt);
				const bool donor = (dm >> lane) & 1u;
				if (donor) wm.donor[drank] = (unsigned char)lane;
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
		const unsigned am = __ballot_sync(FULL, has);
		if (am == 0u && lq == 0) break;

		if (lq >= 32 || am == 0u) {  // (a node step appends at most 64 pairs, 128 with X4: capacity 32 + that)
			// ---- 3. leaf phase: up to 32 triangle pairs ---------------------------------------------------------
			const int k = lq < 32 ? lq : 32;
			lq -= k;
			int verdict = TRI_SEPARATED;
			bool certified = false;
			uint2 pr = make_uint2(0u, 0u);
			unsigned lc = 0, slot = 0;
			if (lane < k) {
				pr = wm.leafq[lq + lane];
				lc = pr.y >> 24;
				slot = wm.cand_slot[lc];
				if (!wm.dead[slot]) {
					const float4 r0 = wm.cand[rec_at(lc, 0)], r1 = wm.cand[rec_at(lc, 1)], r2 = wm.cand[rec_at(lc, 2)],
					             r3 = wm.cand[rec_at(lc, 3)];
					RelF x;
					x.R[0] = r0.x; x.R[1] = r0.y; x.R[2] = r0.z; x.R[3] = r0.w;
					x.R[4] = r1.x; x.R[5] = r1.y; x.R[6] = r1.z; x.R[7] = r1.w;
					x.R[8] = r2.x; x.T[0] = r2.y; x.T[1] = r2.z; x.T[2] = r2.w;
					x.S = r3.w;
					const uint2 base = wm.cand_base[lc];
					const float4 *t1 = p.env.tri32 + 3ull * (base.y + pr.x);
					const float4 *t2 = tris2 + 3ull * (pr.y & 0xffffffu);
					const float4 P1 = __ldg(t1), P2 = __ldg(t1 + 1), P3 = __ldg(t1 + 2);
					const float4 Q1 = __ldg(t2), Q2 = __ldg(t2 + 1), Q3 = __ldg(t2 + 2);
					verdict = tri_filter_f32(P1, P2, P3, Q1, Q2, Q3, x);
					if (verdict == TRI_INTERSECT && !p.always_replay) {
						// path certificate, evaluated from the fp32 record: a larger magnitude
						// only makes it stricter
						const DevModel &m1 = p.env.models[p.objects[p.cand_obj[batch_first + lc]].model];
						const double t1n = (double)((fabsf(x.T[0]) + fabsf(x.T[1]) + fabsf(x.T[2])) * 1.000001f);
						const double mag = t1n + fabs(m1.ctr[0]) + fabs(m1.ctr[1]) + fabs(m1.ctr[2]) +
						                   1.75 * (fabs(m2.ctr[0]) + fabs(m2.ctr[1]) + fabs(m2.ctr[2])) + m1.cmax1 + m2.cmax1 +
						                   m1.esum + m2.esum;
						const double mu = fmin((double)p.env.mu_chain[base.y + pr.x],
						                       (double)p.agent.mu_chain[m2.tri_base + (pr.y & 0xffffffu)]);
						certified = mu >= 1e-5 * mag;
					}
				}
			}
			if (COUNT) ctr[CTR_LEAF] += (unsigned)k;
			if (certified) {  // first contact settles the pose-set
				wm.dead[slot] = 1;
				const uint32_t pose = wm.cand_pose[lc];
				atomicOr(p.verdict_bits + (pose >> 5), 1u << (pose & 31u));
			}
#ifndef MPTCC_NP2_DEFER_ALL
			// Pairs inside the fp32 error band: re-decide them now, in fp64 with FCL's operation order (the 17 axes
			// over 17 lanes), because a contact found here ends the pose-set's traversal.  Measured on C5: leaving
			// them all to exact_kernel costs 11 % more OBB tests and 35 % more triangle tests.
			{
				unsigned um = __ballot_sync(FULL, verdict == TRI_UNCERTAIN);
				while (um) {
					const int from = __ffs(um) - 1;
					um &= um - 1;
					const unsigned c = __shfl_sync(FULL, lc, from), sl = __shfl_sync(FULL, slot, from);
					const uint32_t t1 = __shfl_sync(FULL, pr.x, from), t2 = __shfl_sync(FULL, pr.y, from) & 0xffffffu;
					if (wm.dead[sl]) {  // settled by an earlier pair of this round (warp-uniform: shared memory)
						if (lane == from) verdict = TRI_SEPARATED;
						continue;
					}
					if (COUNT) ctr[CTR_EXACT] += 1;
					const uint2 bs = wm.cand_base[c];
					const bool hit = exact_tri_intersect_warp(p.cand_rel + 12ull * (batch_first + c), p.env.tri64 + 9ull * (bs.y + t1),
					                                          p.agent.tri64 + 9ull * (m2.tri_base + t2), lane);
					if (lane == from) verdict = hit ? TRI_INTERSECT : TRI_SEPARATED;
					if (hit) {  // certificate with the record's magnitudes; uncertified contacts go to exact_kernel (path replay)
						if (lane == from && !p.always_replay) {
							const DevModel &m1 = p.env.models[p.objects[p.cand_obj[batch_first + c]].model];
							const float4 r2 = wm.cand[rec_at(c, 2)];
							const double t1n = (double)((fabsf(r2.y) + fabsf(r2.z) + fabsf(r2.w)) * 1.000001f);
							const double mag = t1n + fabs(m1.ctr[0]) + fabs(m1.ctr[1]) + fabs(m1.ctr[2]) +
							                   1.75 * (fabs(m2.ctr[0]) + fabs(m2.ctr[1]) + fabs(m2.ctr[2])) + m1.cmax1 + m2.cmax1 +
							                   m1.esum + m2.esum;
							const double mu = fmin((double)p.env.mu_chain[bs.y + t1], (double)p.agent.mu_chain[m2.tri_base + t2]);
							if (mu >= 1e-5 * mag) {
								wm.dead[sl] = 1;
								const uint32_t pose = wm.cand_pose[c];
								atomicOr(p.verdict_bits + (pose >> 5), 1u << (pose & 31u));
								verdict = TRI_SEPARATED;  // settled
							}
						}
						__syncwarp();
					}
				}
			}
#endif
			// contacts whose path certificate does not hold (and, with MPTCC_NP2_DEFER_ALL, every pair the filter could
			// not settle) go to the exact pass
			const bool defer = verdict != TRI_SEPARATED && !certified;
			const unsigned dmk = __ballot_sync(FULL, defer);
			if (dmk) {
				unsigned qb = 0;
				if (lane == 0) qb = atomicAdd(p.xq_total, (unsigned)__popc(dmk));
				qb = __shfl_sync(FULL, qb, 0);
				if (defer) {
					const unsigned at = qb + __popc(dmk & lt);
					if (at < p.xq_cap)
						p.xq[at] = make_uint4(batch_first + lc, pr.x, pr.y & 0xffffffu, verdict == TRI_INTERSECT ? 1u : 0u);
					else
						atomicOr(p.overflow, 1u);  // the fused fall-back recomputes the launch
				}
			}
			__syncwarp();
			continue;
		}
		if constexpr (X4) {
			// ---- 4'. node step, both trees expanded: plan = (first environment node, first agent node) with a flag per
			//          side saying whether that side is a sibling pair (expand) or a single node (a leaf, or the root) ----
			constexpr unsigned kExpEnv = 1u << 23, kExpAgent = 1u << 22;
			uint2 lfq[4];
			bool lfv[4] = {false, false, false, false};
			unsigned n_tests = 0;
			if (has) {
				const uint2 pr = cur;
				has = false;
				const unsigned tag = pr.y & 0xff000000u;
				const unsigned nc = pr.y >> 24;
				if (wm.dead[wm.cand_slot[nc]]) {
					top = bot;
				} else {
					const float4 r0 = wm.cand[rec_at(nc, 0)], r1 = wm.cand[rec_at(nc, 1)], r2 = wm.cand[rec_at(nc, 2)];
					const bool xe = (pr.y & kExpEnv) != 0u, xa = (pr.y & kExpAgent) != 0u;
					const uint32_t i1 = pr.x, i2 = pr.y & kPlanIdxMask;
					const unsigned nb1 = wm.cand_base[nc].x;
					Node E[2], D[2];
					if (Q) {
						const uint4 *qnodes1 = p.env.nodeq + 2ull * nb1;
						E[0] = qnode_global(qnodes1 + 2ull * i1);
						E[1] = xe ? qnode_global(qnodes1 + 2ull * i1 + 2) : E[0];
						if (i2 + 1u < ktop) {
							D[0] = qnode_shared(qtop + 2u * i2);
							D[1] = xa ? qnode_shared(qtop + 2u * i2 + 2) : D[0];
						} else {
							D[0] = qnode_global(qnodes2 + 2ull * i2);
							D[1] = xa ? qnode_global(qnodes2 + 2ull * i2 + 2) : D[0];
						}
					} else {
						const float4 *nodes1 = p.env.node32 + 4ull * nb1;
						E[0] = node_global(nodes1 + 4ull * i1);
						E[1] = xe ? node_global(nodes1 + 4ull * i1 + 4) : E[0];
						if (i2 + 1u < ktop) {
							D[0] = node_shared(atop + 4u * i2);
							D[1] = xa ? node_shared(atop + 4u * i2 + 4) : D[0];
						} else {
							D[0] = node_global(nodes2 + 4ull * i2);
							D[1] = xa ? node_global(nodes2 + 4ull * i2 + 4) : D[0];
						}
					}
					// agent nodes into the environment model's frame: W = R [D axes], w = R cD + T
					StepFrame f[2];
#pragma unroll
					for (int j = 0; j < 2; ++j) {
						const Node &S = D[j];
						f[j].W[0] = r0.x * S.v0.x + r0.y * S.v0.y + r0.z * S.v0.z;
						f[j].W[1] = r0.x * S.v1.x + r0.y * S.v1.y + r0.z * S.v1.z;
						f[j].W[2] = r0.x * S.v2.x + r0.y * S.v2.y + r0.z * S.v2.z;
						f[j].W[3] = r0.w * S.v0.x + r1.x * S.v0.y + r1.y * S.v0.z;
						f[j].W[4] = r0.w * S.v1.x + r1.x * S.v1.y + r1.y * S.v1.z;
						f[j].W[5] = r0.w * S.v2.x + r1.x * S.v2.y + r1.y * S.v2.z;
						f[j].W[6] = r1.z * S.v0.x + r1.w * S.v0.y + r2.x * S.v0.z;
						f[j].W[7] = r1.z * S.v1.x + r1.w * S.v1.y + r2.x * S.v1.z;
						f[j].W[8] = r1.z * S.v2.x + r1.w * S.v2.y + r2.x * S.v2.z;
						const float cx = S.v0.w, cy = S.v1.w, cz = S.v2.w;
						f[j].w[0] = r0.x * cx + r0.y * cy + r0.z * cz + r2.y;
						f[j].w[1] = r0.w * cx + r1.x * cy + r1.y * cz + r2.z;
						f[j].w[2] = r1.z * cx + r1.w * cy + r2.x * cz + r2.w;
						f[j].s1 = (fabsf(r2.y) + fabsf(r2.z) + fabsf(r2.w)) + (fabsf(cx) + fabsf(cy) + fabsf(cz));
					}
					uint2 plan[4];
					bool ok[4] = {false, false, false, false};
					float dist[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
					for (int i = 0; i < 2; ++i) {
#pragma unroll
						for (int j = 0; j < 2; ++j) {
							const int k = 2 * i + j;
							if ((i == 1 && !xe) || (j == 1 && !xa)) continue;
							++n_tests;
							float d2;
							if (child_disjoint<Q>(f[j], E[i], D[j].v3, d2)) continue;
							const int linkE = __float_as_int(E[i].v3.w), linkD = __float_as_int(D[j].v3.w);
							if (linkE < 0 && linkD < 0) {
								lfv[k] = true;
								lfq[k] = make_uint2((uint32_t)(-(linkE + 1)), (uint32_t)(-(linkD + 1)) | tag);
								continue;
							}
							// a side that is a leaf stays as it is (single node), the other one is expanded
							const uint32_t nx = linkE >= 0 ? (uint32_t)linkE : i1 + (uint32_t)i;
							const uint32_t ny = linkD >= 0 ? (uint32_t)linkD : i2 + (uint32_t)j;
							plan[k] = make_uint2(nx, ny | (linkE >= 0 ? kExpEnv : 0u) | (linkD >= 0 ? kExpAgent : 0u) | tag);
							ok[k] = true;
							dist[k] = d2;
						}
					}
					// the nearest surviving pair continues, the others are stacked (FCL's order is irrelevant to the verdict)
					int best = -1;
#pragma unroll
					for (int k = 0; k < 4; ++k)
						if (ok[k] && (best < 0 || (p.near_first && dist[k] < dist[best]))) best = k;
					if (best >= 0) {
#pragma unroll
						for (int k = 3; k >= 0; --k)
							if (ok[k] && k != best) stk[top++ & M] = plan[k];
						cur = plan[best];
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
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				const unsigned lm = __ballot_sync(FULL, lfv[k]);
				if (lfv[k]) wm.leafq[lq + __popc(lm & lt)] = lfq[k];
				lq += __popc(lm);
			}
			__syncwarp();
			continue;
		}
		// ---- 4. node step: both children of the descended tree against the stationary node of the other ----------
		bool leaf0 = false, leaf1 = false;
		uint2 lf0 = make_uint2(0u, 0u), lf1 = make_uint2(0u, 0u);
		unsigned n_tests = 0;
		if (has) {
			const uint2 pr = cur;
			has = false;
			const unsigned tag = pr.y & 0xff000000u;
			const unsigned nc = pr.y >> 24;
			if (wm.dead[wm.cand_slot[nc]]) {
				top = bot;  // this lane's stack holds one candidate only, and its pose-set already collides
			} else {
				const float4 r0 = wm.cand[rec_at(nc, 0)], r1 = wm.cand[rec_at(nc, 1)], r2 = wm.cand[rec_at(nc, 2)],
				             r3 = wm.cand[rec_at(nc, 3)];
				const bool single = (pr.y & kPlanSingle) != 0u;
				const bool env_side = single || (pr.y & kPlanEnv) != 0u;  // the children are environment nodes
				const uint32_t i2base = pr.y & kPlanIdxMask;
				const unsigned nb1 = wm.cand_base[nc].x;
				Node S, C0, C1;
				if (Q) {
					const uint4 *qnodes1 = p.env.nodeq + 2ull * nb1;
					if (env_side) {
						S = i2base < ktop ? qnode_shared(qtop + 2u * i2base) : qnode_global(qnodes2 + 2ull * i2base);
						C0 = qnode_global(qnodes1 + 2ull * pr.x);
						C1 = single ? C0 : qnode_global(qnodes1 + 2ull * pr.x + 2);
					} else {
						S = qnode_global(qnodes1 + 2ull * pr.x);
						if (i2base + 1u < ktop) {
							C0 = qnode_shared(qtop + 2u * i2base);
							C1 = qnode_shared(qtop + 2u * i2base + 2);
						} else {
							C0 = qnode_global(qnodes2 + 2ull * i2base);
							C1 = qnode_global(qnodes2 + 2ull * i2base + 2);
						}
					}
				} else {
					const float4 *nodes1 = p.env.node32 + 4ull * nb1;
					if (env_side) {
						S = i2base < ktop ? node_shared(atop + 4u * i2base) : node_global(nodes2 + 4ull * i2base);
						C0 = node_global(nodes1 + 4ull * pr.x);
						C1 = single ? C0 : node_global(nodes1 + 4ull * pr.x + 4);
					} else {
						S = node_global(nodes1 + 4ull * pr.x);
						if (i2base + 1u < ktop) {
							C0 = node_shared(atop + 4u * i2base);
							C1 = node_shared(atop + 4u * i2base + 4);
						} else {
							C0 = node_global(nodes2 + 4ull * i2base);
							C1 = node_global(nodes2 + 4ull * i2base + 4);
						}
					}
				}
				// (Ry, Ty): the relative transform when the children are environment nodes, its inverse otherwise
				StepFrame f;
				{
					const float y00 = r0.x, y11 = r1.x, y22 = r2.x;
					const float y01 = env_side ? r0.y : r0.w, y10 = env_side ? r0.w : r0.y;
					const float y02 = env_side ? r0.z : r1.z, y20 = env_side ? r1.z : r0.z;
					const float y12 = env_side ? r1.y : r1.w, y21 = env_side ? r1.w : r1.y;
					const float ty0 = env_side ? r2.y : r3.x, ty1 = env_side ? r2.z : r3.y, ty2 = env_side ? r2.w : r3.z;
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
				}
				const int linkS = __float_as_int(S.v3.w);
				const float sS = S.v3.x * S.v3.x + S.v3.y * S.v3.y + S.v3.z * S.v3.z;
				uint2 plan0 = make_uint2(0u, 0u), plan1 = make_uint2(0u, 0u);
				bool ok0 = false, ok1 = false;
				float d2_0 = 0.f, d2_1 = 0.f;
#pragma unroll
				for (int k = 0; k < 2; ++k) {
					if (k == 1 && single) break;
					const Node &C = k == 0 ? C0 : C1;
					++n_tests;
					float d2;
					if (child_disjoint<Q>(f, C, S.v3, d2)) continue;
					if (k == 0) d2_0 = d2; else d2_1 = d2;
					const int linkC = __float_as_int(C.v3.w);
					const float sC = C.v3.x * C.v3.x + C.v3.y * C.v3.y + C.v3.z * C.v3.z;
					// back to (environment, agent) roles
					const int link1 = env_side ? linkC : linkS, link2 = env_side ? linkS : linkC;
					const uint32_t i1 = pr.x + ((env_side && k) ? 1u : 0u);
					const uint32_t i2 = i2base + ((!env_side && k) ? 1u : 0u);
					const bool l1 = link1 < 0, l2 = link2 < 0;
					if (l1 && l2) {
						const uint2 lf = make_uint2((uint32_t)(-(link1 + 1)), (uint32_t)(-(link2 + 1)) | tag);
						if (k == 0) { leaf0 = true; lf0 = lf; } else { leaf1 = true; lf1 = lf; }
						continue;
					}
					// firstOverSecond: descend the environment tree iff l2 || (!l1 && size1 > size2)
					const float s1 = env_side ? sC : sS, s2 = env_side ? sS : sC;
					const bool first = l2 || (!l1 && s1 > s2);
					const uint2 pl = first ? make_uint2((uint32_t)link1, i2 | kPlanEnv | tag) : make_uint2(i1, (uint32_t)link2 | tag);
					if (k == 0) { ok0 = true; plan0 = pl; } else { ok1 = true; plan1 = pl; }
				}
				if (ok0 && ok1) {
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
		const unsigned lm = __ballot_sync(FULL, leaf1);
		if (leaf1) wm.leafq[lq + __popc(lm & lt)] = lf1;
		lq += __popc(lm);
		__syncwarp();
	}
}

#ifdef MPTCC_NP2_MAXREG
#define NP2_BOUNDS __maxnreg__(MPTCC_NP2_MAXREG)
#else
#define NP2_BOUNDS __launch_bounds__(kNp2Warps * 32, MPTCC_NP2_CTAS)
#endif

template <bool COUNT, int BATCH, bool Q, bool X4>
__global__ void NP2_BOUNDS narrowphase2_kernel(const __grid_constant__ KernelParams p) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	float4 *atop = reinterpret_cast<float4 *>(smem_raw);
	const DevModel &m2 = p.agent.models[p.link_model[0]];

	// ---- TMA bulk copy of the agent tree's top levels into shared memory (one thread issues, all wait) ---------
	const unsigned top_bytes = p.agent_top * (Q ? 32u : 64u);
	const unsigned mbar = (unsigned)__cvta_generic_to_shared(smem_raw + kTopBytes);
	if (top_bytes != 0u) {
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(mbar), "r"(1u) : "memory");
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			const void *src = Q ? static_cast<const void *>(p.agent.nodeq + 2ull * m2.node_base)
			                    : static_cast<const void *>(p.agent.node32 + 4ull * m2.node_base);
			const unsigned dst = (unsigned)__cvta_generic_to_shared(atop);
			asm volatile("mbarrier.arrive.expect_tx.relaxed.cta.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(top_bytes) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
			             "l"(src), "r"(top_bytes), "r"(mbar)
			             : "memory");
		}
	}

	unsigned char *base = smem_raw + kTopRegion + (size_t)warp * np2_warp_smem_bytes(BATCH, X4);
	Np2Mem wm;
	wm.leafq = reinterpret_cast<uint2 *>(base);
	base += (X4 ? kLeafQueueCapX4 : (unsigned)kLeafQueueCap) * 8;
	wm.cand = reinterpret_cast<float4 *>(base);
	base += BATCH * 64;
	wm.cand_base = reinterpret_cast<uint2 *>(base);
	base += BATCH * 8;
	wm.cand_pose = reinterpret_cast<uint32_t *>(base);
	base += BATCH * 4;
	wm.cand_slot = base;
	base += BATCH;
	wm.dead = base;
	base += BATCH;
	wm.donor = base;

	// overflowed list / non-rigid pose: the whole launch belongs to the fused fall-back kernel
	const bool skip = *p.overflow != 0u;
	if (top_bytes != 0u) {  // wait for the staged block (every thread observes the phase flip)
		unsigned done = 0;
		while (!done)
			asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
			             : "=r"(done)
			             : "r"(mbar), "r"(0u)
			             : "memory");
	}
	if (skip) return;
	const unsigned total = *p.cand_total;
	unsigned long long ctr[CTR_N];
#pragma unroll
	for (int i = 0; i < CTR_N; ++i) ctr[i] = 0;

	while (true) {
		unsigned first = 0;
		if (lane == 0) first = atomicAdd(p.work_counter2, (unsigned)BATCH);
		first = __shfl_sync(FULL, first, 0);
		if (first >= total) break;
		const int n = (int)(total - first < (unsigned)BATCH ? total - first : (unsigned)BATCH);
		// ---- load the batch ---------------------------------------------------------------------------------------
		const float4 *src_rec = p.cand_rec + 4ull * first;
#ifdef MPTCC_LIST_EF
		for (int i = lane; i < 2 * n; i += 32) {
			const F8 v = ldg256_evict_first(src_rec + 2 * i);
			wm.cand[rec_at((unsigned)i >> 1, 2u * ((unsigned)i & 1u))] = v.a;
			wm.cand[rec_at((unsigned)i >> 1, 2u * ((unsigned)i & 1u) + 1u)] = v.b;
		}
#else
		for (int i = lane; i < 4 * n; i += 32) wm.cand[rec_at((unsigned)i >> 2, (unsigned)i & 3u)] = src_rec[i];
#endif
		for (int i = lane; i < n; i += 32) {
			wm.cand_base[i] = p.cand_base[first + i];
			wm.cand_pose[i] = p.cand_pose[first + i];
		}
		__syncwarp();
		// ---- pose slots: candidates of one pose-set are adjacent; the slot of a run is the index of its first
		//      candidate; dead[slot] starts from the verdict bit (another batch may have settled the pose-set) ------
		unsigned carry = 0;
		for (int g = 0; g < (n + 31) / 32; ++g) {
			const int i = 32 * g + lane;
			const bool valid = i < n;
			const uint32_t pose = valid ? wm.cand_pose[i] : 0u;
			const bool start = valid && (i == 0 || wm.cand_pose[i - 1] != pose);
			const unsigned sm = __ballot_sync(FULL, start);
			const unsigned upto = sm & (0xffffffffu >> (31 - lane));
			const unsigned slot = upto ? (unsigned)(32 * g + 31 - __clz((int)upto)) : carry;
			if (valid) wm.cand_slot[i] = (unsigned char)slot;
			if (start) wm.dead[i] = (unsigned char)((p.verdict_bits[pose >> 5] >> (pose & 31u)) & 1u);
			if (sm) carry = (unsigned)(32 * g + 31 - __clz((int)sm));
		}
		__syncwarp();
		traverse2<COUNT, BATCH, Q, X4>(p, wm, atop, m2, n, first, lane, ctr);
		__syncwarp();
	}
	if (COUNT && lane == 0) {
#pragma unroll
		for (int i = 0; i < CTR_N; ++i)
			if (ctr[i]) atomicAdd(p.counters + i, ctr[i]);
	}
}

// ---- exact pass: one warp per queued leaf pair -----------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(256) exact_kernel(const __grid_constant__ KernelParams p) {
	if (*p.overflow != 0u) return;
	const int lane = threadIdx.x & 31;
	const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const unsigned n_warps = (gridDim.x * blockDim.x) >> 5;
	const unsigned total = min(*p.xq_total, p.xq_cap);
	const DevModel &m2 = p.agent.models[p.link_model[0]];
	unsigned long long n_exact = 0, n_path = 0;
	for (unsigned i = warp; i < total; i += n_warps) {
		const uint4 e = p.xq[i];
		const uint32_t pose = p.cand_pose[e.x];
		if ((p.verdict_bits[pose >> 5] >> (pose & 31u)) & 1u) continue;  // settled meanwhile
		const DevObject &ob = p.objects[p.cand_obj[e.x]];
		const DevModel &m1 = p.env.models[ob.model];
		const double *rel = p.cand_rel + 12ull * e.x;
		const bool sure = e.w != 0u && !p.always_replay;
		if (COUNT && !sure) ++n_exact;
		const bool hit = sure ? true
		                      : exact_tri_intersect_warp(rel, p.env.tri64 + 9ull * (m1.tri_base + e.y),
		                                                 p.agent.tri64 + 9ull * (m2.tri_base + e.z), lane);
		if (!hit) continue;
		const double mag = fabs(rel[9]) + fabs(rel[10]) + fabs(rel[11]) + m1.cmax1 + m2.cmax1 + m1.esum + m2.esum;
		const double mu = fmin((double)p.env.mu_chain[m1.tri_base + e.y], (double)p.agent.mu_chain[m2.tri_base + e.z]);
		bool reached = !p.always_replay && mu >= 1e-5 * mag;
		if (!reached) {
			if (COUNT) ++n_path;
			reached = exact_path_reaches(p, m1, m2, rel, e.y, e.z, lane);
		}
		if (reached && lane == 0) atomicOr(p.verdict_bits + (pose >> 5), 1u << (pose & 31u));
	}
	if (COUNT && lane == 0) {
		if (n_exact) atomicAdd(p.counters + CTR_EXACT, n_exact);
		if (n_path) atomicAdd(p.counters + CTR_PATH, n_path);
	}
}

template <class K>
int launch_np2(K kernel, const KernelParams &p, int sm_count, size_t smem, void *stream) {
	int per_sm = 0;
	const cudaError_t e = prepare_kernel(reinterpret_cast<const void *>(kernel), kNp2Warps * 32, smem, &per_sm);
	if (e != cudaSuccess) return (int)e;
	if (per_sm < 1) per_sm = 1;
	kernel<<<sm_count * per_sm, kNp2Warps * 32, smem, static_cast<cudaStream_t>(stream)>>>(p);
	return (int)cudaGetLastError();
}

}  // namespace

int launch_narrowphase2(const KernelParams &p, int sm_count, bool count, bool long_batches, void *stream) {
#define NP2_LAUNCH_X(B, X)                                                                                             \
	{                                                                                                                  \
		const size_t smem = kTopRegion + (size_t)kNp2Warps * np2_warp_smem_bytes(B, X);                                \
		if (p.qnode)                                                                                                   \
			return count ? launch_np2(narrowphase2_kernel<true, B, true, X>, p, sm_count, smem, stream)                \
			             : launch_np2(narrowphase2_kernel<false, B, true, X>, p, sm_count, smem, stream);              \
		return count ? launch_np2(narrowphase2_kernel<true, B, false, X>, p, sm_count, smem, stream)                   \
		             : launch_np2(narrowphase2_kernel<false, B, false, X>, p, sm_count, smem, stream);                 \
	}
#define NP2_LAUNCH(B)                   \
	{                                   \
		if (p.expand4) NP2_LAUNCH_X(B, true) \
		NP2_LAUNCH_X(B, false)          \
	}
	if (long_batches && kNp2BatchLong != kNp2Batch) NP2_LAUNCH(kNp2BatchLong)
	NP2_LAUNCH(kNp2Batch)
#undef NP2_LAUNCH_X
#undef NP2_LAUNCH
}

int launch_exact(const KernelParams &p, int sm_count, bool count, void *stream) {
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	if (count)
		exact_kernel<true><<<sm_count * 2, 256, 0, s>>>(p);
	else
		exact_kernel<false><<<sm_count * 2, 256, 0, s>>>(p);
	return (int)cudaGetLastError();
}

}  // namespace mptcc
