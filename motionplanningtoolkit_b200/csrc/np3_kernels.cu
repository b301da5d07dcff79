// np3_kernels.cu -- narrowphase of the split pipeline, third generation: STREAMING lanes.
//
// The first two generations hand a warp a batch of 96/128 candidates and let its lanes steal pending sub-trees from
// each other once the batch has no unstarted candidate left.  A simulation of that schedule on the real geometry
// (tools/research/warp_sim.c) shows what it costs on the benchmark scene: every batch ends in a tail in which the
// lanes explore, speculatively and in parallel, the sub-trees of the few long candidates -- 57-61 OBB tests per
// pose-set instead of the 50 of a schedule without batch ends, and 26-29 busy lanes per step instead of 31.
// Here a lane that runs out of work simply starts the next candidate of the global list:
//
//   * every lane owns a ring of three shared-memory slots: the record of its CURRENT candidate (read back with four
//     conflict-free LDS.128 per step -- keeping it in registers was measured: the kernel then needs 164 registers and
//     spills at 128) and the records of its next TWO candidates, prefetched with cp.async while the lane is still
//     busy, so starting a candidate costs no round trip to the list; no shared candidate table, no batch;
//   * list indices are reserved in chunks of 64 with one atomicAdd per chunk;
//   * lanes only steal from each other once the list is exhausted (the tail of the launch);
//   * a contact is published with atomicOr and broadcast inside the warp by pose-set id (lanes working on the same
//     pose-set drop their work; a small ring of recently settled pose-sets stops prefetched candidates and queued
//     leaf pairs of that pose-set);
//   * the whole agent tree (breadth-first numbering, as many nodes as fit beside the per-warp state: 2 590 of the
//     blimp's 2 709) is staged in shared memory with TMA bulk copies, in a bank-swizzled layout prepared on the host,
//     so that only environment nodes are gathered from L1/L2.
// The SAT (child-frame, shared stationary half), the triangle filter, the inline fp64 re-decide of uncertain pairs and
// the exact-pass queue are those of np2_kernels.cu.
#include "collide_device.cuh"
#include "np_sat.cuh"
#include "launch_cache.hpp"
#ifdef MPTCC_NP3_TRACE
#include <algorithm>
#include <cstdio>
#include <vector>
#endif

namespace mptcc {
namespace {

#ifndef MPTCC_NP3_WARPS
#define MPTCC_NP3_WARPS 16
#endif
#ifndef MPTCC_NP3_CHUNK
#define MPTCC_NP3_CHUNK 64
#endif
#ifndef MPTCC_NP3_LEAF_AT
#define MPTCC_NP3_LEAF_AT 32
#endif
constexpr int kNp3Warps = MPTCC_NP3_WARPS;
constexpr unsigned kNp3Chunk = MPTCC_NP3_CHUNK;
constexpr int kNp3LeafAt = MPTCC_NP3_LEAF_AT;  // queued leaf pairs that trigger a leaf phase
constexpr int kNp3LeafCap = kNp3LeafAt + 64;   // a node step appends at most 64
constexpr unsigned kNp3SlotBytes = 80u;        // 4 float4 record + (node_base, tri_base, pose-set, -)
constexpr unsigned kNp3Depth = 2u;             // candidates prefetched per lane, consumed in order
constexpr unsigned kNp3Ring = kNp3Depth + 1u;  // slots per lane: the current candidate + the prefetched ones
#ifndef MPTCC_NP3_STACK_S
#define MPTCC_NP3_STACK_S 12
#endif
constexpr unsigned kNp3StackS = MPTCC_NP3_STACK_S;  // per-lane stack entries kept in shared memory (deeper ones: local memory)
constexpr unsigned kNp3WarpBytes = 32u * kNp3Ring * kNp3SlotBytes + (unsigned)kNp3LeafCap * 16u + 32u + 32u + kNp3StackS * 256u;
constexpr unsigned kNp3TopOffset = 128u;       // the mbarrier lives in front of the staged tree
constexpr unsigned kPlan3Env = 1u << 31, kPlan3Single = 1u << 30, kPlan3Idx = (1u << 30) - 1u;

// Node i of the staged agent tree: part k (of 4 float4) sits at position (k + (i >> 1)) & 3 of its 64 bytes, so that
// lanes fetching the same part of different nodes spread over all eight 16-byte bank groups.
__device__ __forceinline__ Node node_shared_swz(const float4 *top, unsigned i) {
	const float4 *b = top + 4u * i;
	const unsigned r = i >> 1;
	return Node{b[r & 3u], b[(r + 1u) & 3u], b[(r + 2u) & 3u], b[(r + 3u) & 3u]};
}

__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// every group but the most recent one has landed
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

__device__ __forceinline__ bool pose_settled(const uint32_t *hits, uint32_t pose) {
	const uint4 a = *reinterpret_cast<const uint4 *>(hits), b = *reinterpret_cast<const uint4 *>(hits + 4);
	return a.x == pose || a.y == pose || a.z == pose || a.w == pose || b.x == pose || b.y == pose || b.z == pose || b.w == pose;
}

#ifndef MPTCC_NP3_CTAS
#define MPTCC_NP3_CTAS 1
#endif
#ifdef MPTCC_NP3_MAXREG
#define NP3_BOUNDS __maxnreg__(MPTCC_NP3_MAXREG)
#else
#define NP3_BOUNDS __launch_bounds__(kNp3Warps * 32, MPTCC_NP3_CTAS)
#endif

#ifdef MPTCC_NP3_TRACE
// Research build (tools/build_variants.sh ... "-DMPTCC_NP3_TRACE"): per-warp time stamps (start, list exhausted, exit)
// and node steps / busy lanes before and after the list ran out, printed by launch_narrowphase3 every 8th launch.
__device__ unsigned long long np3_trace[8192 * 5];
__device__ __forceinline__ unsigned long long gtime() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}
#endif

template <bool COUNT, bool TOP, bool Q>
__global__ void NP3_BOUNDS narrowphase3_kernel(const __grid_constant__ KernelParams p) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const unsigned lt = (1u << lane) - 1u;
	const DevModel &m2 = p.agent.models[p.link_model[0]];
	const unsigned ktop = TOP ? p.agent_top : 0u;
	const unsigned top_bytes = ktop * 64u;
	const float4 *atop = reinterpret_cast<const float4 *>(smem_raw + kNp3TopOffset);

	// ---- TMA bulk copies of the (swizzled) agent tree into shared memory: one thread issues, all wait ------------
	const unsigned mbar = (unsigned)__cvta_generic_to_shared(smem_raw);
	if (top_bytes != 0u) {
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(mbar), "r"(1u) : "memory");
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.arrive.expect_tx.relaxed.cta.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(top_bytes) : "memory");
			const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_raw + kNp3TopOffset);
			const char *src = reinterpret_cast<const char *>(p.agent_swz);
			for (unsigned off = 0; off < top_bytes; off += 32768u) {
				const unsigned len = top_bytes - off < 32768u ? top_bytes - off : 32768u;
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + off),
				             "l"(src + off), "r"(len), "r"(mbar)
				             : "memory");
			}
		}
	}

	unsigned char *wbase = smem_raw + kNp3TopOffset + ((top_bytes + 127u) & ~127u) + (size_t)warp * kNp3WarpBytes;
	// record slots of the warp, [kNp3Ring][32 lanes][5 float4]: slot r of lane l starts 5 (32 r + l) float4 in; 80-byte
	// stride over the lanes and a multiple of 128 bytes between ring positions, so lanes reading part k of their
	// current slot hit eight different 16-byte bank groups per quarter-warp whatever their ring positions are
	float4 *wslots = reinterpret_cast<float4 *>(wbase);
#define NP3_SLOT(l, r) (wslots + 5u * (32u * (r) + (unsigned)(l)))
	uint4 *leafq = reinterpret_cast<uint4 *>(wbase + 32u * kNp3Ring * kNp3SlotBytes);  // [kNp3LeafCap] env prim, agent prim, list index, pose-set
	uint32_t *hits = reinterpret_cast<uint32_t *>(wbase + 32u * kNp3Ring * kNp3SlotBytes + (unsigned)kNp3LeafCap * 16u);  // [8]
	unsigned char *donor = reinterpret_cast<unsigned char *>(hits + 8);       // [32]
	// The pending plans of a lane: the node gathers keep evicting a local-memory stack from L1 (87 % of its loads
	// missed), so the entries every descent uses live in shared memory, word-interleaved over the lanes.
	uint32_t *sstk = reinterpret_cast<uint32_t *>(donor + 32) + lane;         // [kNp3StackS][2][32]
#define NP3_PUSH(v)                                                  \
	{                                                                \
		const uint2 v_ = (v);                                        \
		if (top < kNp3StackS) {                                      \
			sstk[64u * top] = v_.x;                                  \
			sstk[64u * top + 32u] = v_.y;                            \
		} else {                                                     \
			stk[top - kNp3StackS] = v_;                              \
		}                                                            \
		++top;                                                       \
	}
#define NP3_POP(dst)                                                 \
	{                                                                \
		--top;                                                       \
		if (top < kNp3StackS) dst = make_uint2(sstk[64u * top], sstk[64u * top + 32u]); \
		else dst = stk[top - kNp3StackS];                            \
	}

	const bool skip = *p.overflow != 0u;  // overflowed list / non-rigid pose: the launch belongs to the fused fall-back
	if (top_bytes != 0u) {
		unsigned done = 0;
		while (!done)
			asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
			             : "=r"(done)
			             : "r"(mbar), "r"(0u)
			             : "memory");
	}
	if (skip) return;
	const unsigned total = *p.cand_total;
	if (lane < 8) hits[lane] = 0xffffffffu;
#ifdef MPTCC_NP3_TRACE
	unsigned long long *tr = np3_trace + 4ull * ((blockIdx.x * (unsigned)kNp3Warps + (unsigned)warp) & 8191u);
	if (lane == 0) {
		tr[0] = gtime();
		tr[1] = 0;
	}
	unsigned long long fin_steps = 0, run_steps = 0;
#endif
	__syncwarp();

	const float4 *__restrict__ nodes2 = p.agent.node32 + 4ull * m2.node_base;
	const float4 *__restrict__ tris2 = p.agent.tri32 + 3ull * m2.tri_base;
	unsigned long long ctr[CTR_N];
#pragma unroll
	for (int i = 0; i < CTR_N; ++i) ctr[i] = 0;

	// ---- lane state ----------------------------------------------------------------------------------------------
	unsigned nb1 = 0, ci = 0, pose = 0xffffffffu;  // env node_base, list index, pose-set of the current candidate
	unsigned rd = 0;                     // ring slot of the current candidate; the prefetched ones follow it
	unsigned pf = 0, young = 0;          // prefetched candidates, how many of them were issued in the most recent
	                                     // cp.async group (not consumable yet)
	uint2 stk[kLaneStackCap];
	unsigned top = 0;
	uint2 cur = make_uint2(0u, 0u);
	bool has = false;
	unsigned qhead = 0, qend = 0, hit_w = 0;
	unsigned seen = 0;  // end of this warp's last reservation: how far the list had been handed out at that time
	const unsigned n_warps_all = gridDim.x * (unsigned)kNp3Warps;
	bool exhausted = false;
	int lq = 0;

	while (true) {
		if (!has && top != 0u) {
			NP3_POP(cur)
			has = true;
		}
		unsigned im = __ballot_sync(FULL, !has);
		if (im) {
			const unsigned cm = __ballot_sync(FULL, pf > young);  // lanes holding a prefetched candidate that has landed
			if (cm) {
				cp_async_wait_but_one();
				__syncwarp();
				if (exhausted) {
					// tail of the launch: prefetched candidates of busy lanes go to idle lanes that have none
					const unsigned want = im & __ballot_sync(FULL, pf == 0u), give = cm & ~im;
					if (want && give) {
						const int grank = __popc(give & lt);
						if ((give >> lane) & 1u) donor[grank] = (unsigned char)lane;
						__syncwarp();
						const int r = __popc(want & lt);
						const bool taker = ((want >> lane) & 1u) && r < __popc(give);
						const int src = taker ? (int)donor[r] : lane;
						const unsigned srd = __shfl_sync(FULL, rd, src);
						const bool given = ((give >> lane) & 1u) && grank < __popc(want);
						if (taker) {
							// copy the donor's oldest prefetched record into my next slot, as if I had prefetched it
							const float4 *from = NP3_SLOT(src, (srd + 1u) % kNp3Ring);
							float4 *to = NP3_SLOT(lane, (rd + 1u) % kNp3Ring);
#pragma unroll
							for (int k = 0; k < 5; ++k) to[k] = from[k];
							pf = 1u;
						}
						__syncwarp();
						if (given) {
							// the donor's remaining prefetched record (if any) moves up: rotate by making the given slot the
							// "current" one is not possible while the lane is busy, so swap the two prefetched slots instead
							if (pf == 2u) {
								float4 *a = NP3_SLOT(lane, (rd + 1u) % kNp3Ring);
								const float4 *bsl = NP3_SLOT(lane, (rd + 2u) % kNp3Ring);
#pragma unroll
								for (int k = 0; k < 5; ++k) a[k] = bsl[k];
							}
							--pf;
						}
						__syncwarp();
					}
				}
				if (!has && pf > young) {
					rd = (rd + 1u) % kNp3Ring;
					--pf;
					const float4 s4 = NP3_SLOT(lane, rd)[4];
					ci = __float_as_uint(s4.w);
					pose = __float_as_uint(s4.z);
					nb1 = __float_as_uint(s4.x);
					if (!pose_settled(hits, pose)) {
						cur = make_uint2(0u, kPlan3Single | kPlan3Env);  // root pair: (environment root, agent root)
						has = true;
					}
				}
				im = __ballot_sync(FULL, !has);
			}
		}
		// ---- prefetch: every lane keeps kNp3Depth candidates in flight ---------------------------------------------
		young = 0u;
		if (im && !exhausted) {
			// Guided self-scheduling: list indices are reserved in chunks that shrink as the list drains (a quarter of an
			// even share of what was left at this warp's last reservation, 8..kNp3Chunk), and near the end a lane keeps one
			// candidate in flight instead of two -- short launches (the 2^18-pose pieces of the host-buffer path give a
			// warp ~150 candidates) would otherwise end with a few warps holding the last hundreds of candidates.
			const unsigned left = total > seen ? total - seen : 0u;
			const unsigned depth = left > n_warps_all * 64u ? kNp3Depth : 1u;
			unsigned need = __ballot_sync(FULL, pf < depth);
			while (need) {
				if (qhead == qend) {
					unsigned want = left / (4u * n_warps_all);
					want = want < 8u ? 8u : (want > kNp3Chunk ? kNp3Chunk : want);
					unsigned first = 0;
					if (lane == 0) first = atomicAdd(p.work_counter2, want);
					first = __shfl_sync(FULL, first, 0);
					seen = first + want;
					if (first >= total) {
						exhausted = true;
#ifdef MPTCC_NP3_TRACE
						if (lane == 0) tr[1] = gtime();
#endif
						break;
					}
					qhead = first;
					qend = total - first < want ? total : first + want;
				}
				const unsigned avail = qend - qhead;
				const unsigned rank = (unsigned)__popc(need & lt);
				if (((need >> lane) & 1u) && rank < avail) {
					const unsigned idx = qhead + rank;
					float4 *s = NP3_SLOT(lane, (rd + pf + 1u) % kNp3Ring);
					const float4 *g = p.cand_rec + 4ull * idx;
					cp_async16(s, g);
					cp_async16(s + 1, g + 1);
					cp_async16(s + 2, g + 2);
					cp_async16(s + 3, g + 3);
					cp_async8(s + 4, p.cand_base + idx);
					cp_async4(reinterpret_cast<char *>(s + 4) + 8, p.cand_pose + idx);
					reinterpret_cast<unsigned *>(s + 4)[3] = idx;
					++pf;
					++young;
				}
				const unsigned nneed = (unsigned)__popc(need);
				qhead += nneed < avail ? nneed : avail;
				need = __ballot_sync(FULL, pf < depth);
			}
		}
		cp_async_commit();
		if (im) {
			// ---- tail: idle lanes steal the top pending plan of a busy lane (with its candidate) ---------------------
			if (exhausted && !p.no_steal) {
				const unsigned dm = __ballot_sync(FULL, has && top != 0u);
				const unsigned tm = im & ~__ballot_sync(FULL, pf != 0u);  // idle lanes without a prefetched candidate of their own
				if (dm && tm) {
					const int n_thieves = __popc(tm), n_donors = __popc(dm);
					const int drank = __popc(dm & lt);
					const bool is_donor = (dm >> lane) & 1u;
					if (is_donor) donor[drank] = (unsigned char)lane;
					__syncwarp();
					const int r = __popc(tm & lt);
					const bool thief = ((tm >> lane) & 1u) && r < n_donors;
					const int from = thief ? (int)donor[r] : lane;
					uint2 e = make_uint2(0u, 0u);
					if (is_donor && drank < n_thieves) NP3_POP(e)
					e.x = __shfl_sync(FULL, e.x, from);
					e.y = __shfl_sync(FULL, e.y, from);
					const unsigned frd = __shfl_sync(FULL, rd, from);
					const unsigned fnb1 = __shfl_sync(FULL, nb1, from), fci = __shfl_sync(FULL, ci, from), fpose = __shfl_sync(FULL, pose, from);
					if (thief) {
						const float4 *src = NP3_SLOT(from, frd);
						float4 *to = NP3_SLOT(lane, rd);
#pragma unroll
						for (int k = 0; k < 5; ++k) to[k] = src[k];  // (the record and node_base, tri_base, pose-set, list index)
						nb1 = fnb1;
						ci = fci;
						pose = fpose;
					}
					if (thief) {
						cur = e;
						has = true;
					}
					__syncwarp();
				}
			}
		}
		const unsigned am = __ballot_sync(FULL, has);
		if (am == 0u && lq == 0) {
			if (exhausted && __ballot_sync(FULL, pf != 0u) == 0u) break;
			continue;
		}

		if (lq >= kNp3LeafAt || am == 0u) {
			// ---- leaf phase: up to 32 triangle pairs ------------------------------------------------------------------
			const int k = lq < 32 ? lq : 32;
			lq -= k;
			int verdict = TRI_SEPARATED;
			bool certified = false;
			uint4 e = make_uint4(0u, 0u, 0u, 0xfffffffeu);  // (environment triangle in the pack, agent triangle, list index, pose-set)
			float mu = 0.f, cert = 0.f;
			if (lane < k) {
				e = leafq[lq + lane];
				if (!pose_settled(hits, e.w)) {
					// every load of the phase is issued before the first use: one memory round trip, not three
					const float4 *rp = p.cand_rec + 4ull * e.z;
					const F8 ra = ldg256(rp), rb = ldg256(rp + 2);
					const float4 *t1 = p.env.tri32 + 3ull * e.x;
					const float4 *t2 = tris2 + 3ull * e.y;
					const float4 P1 = __ldg(t1), P2 = __ldg(t1 + 1), P3 = __ldg(t1 + 2);
					const float4 Q1 = __ldg(t2), Q2 = __ldg(t2 + 1), Q3 = __ldg(t2 + 2);
					// path certificate: the smaller mu_chain of the two triangles against the candidate's
					// threshold from the broadphase kernel
					mu = fminf(__ldg(p.env.mu_chain + e.x), __ldg(p.agent.mu_chain + m2.tri_base + e.y));
					cert = __ldg(p.cand_cert + e.z);
					RelF x;
					x.R[0] = ra.a.x; x.R[1] = ra.a.y; x.R[2] = ra.a.z; x.R[3] = ra.a.w;
					x.R[4] = ra.b.x; x.R[5] = ra.b.y; x.R[6] = ra.b.z; x.R[7] = ra.b.w;
					x.R[8] = rb.a.x; x.T[0] = rb.a.y; x.T[1] = rb.a.z; x.T[2] = rb.a.w;
					x.S = rb.b.w;
					verdict = tri_filter_f32(P1, P2, P3, Q1, Q2, Q3, x);
					certified = verdict == TRI_INTERSECT && !p.always_replay && mu >= cert;
				}
			}
			if (COUNT) ctr[CTR_LEAF] += (unsigned)k;
			if (certified) atomicOr(p.verdict_bits + (e.w >> 5), 1u << (e.w & 31u));  // first contact settles the pose-set
			unsigned hm = __ballot_sync(FULL, certified);
#ifndef MPTCC_NP3_DEFER_ALL
			// Pairs inside the fp32 error band: re-decide them now, in fp64 with FCL's operation order (the 17 axes over
			// 17 lanes), because a contact found here ends the pose-set's traversal.
			{
				unsigned um = __ballot_sync(FULL, verdict == TRI_UNCERTAIN);
				while (um) {
					const int from = __ffs(um) - 1;
					um &= um - 1;
					const uint32_t c = __shfl_sync(FULL, e.z, from), ps = __shfl_sync(FULL, e.w, from);
					const uint32_t t1 = __shfl_sync(FULL, e.x, from), t2 = __shfl_sync(FULL, e.y, from);
					// settled by an earlier pair of this round?  (warp-uniform)
					unsigned same = __ballot_sync(FULL, ((hm >> lane) & 1u) && e.w == ps);
					if (same) {
						if (lane == from) verdict = TRI_SEPARATED;
						continue;
					}
					if (COUNT) ctr[CTR_EXACT] += 1;
					const bool hit = exact_tri_intersect_warp(p.cand_rel + 12ull * c, p.env.tri64 + 9ull * t1,
					                                          p.agent.tri64 + 9ull * (m2.tri_base + t2), lane);
					if (lane == from) verdict = hit ? TRI_INTERSECT : TRI_SEPARATED;
					if (hit) {  // certificate with the record's magnitudes; uncertified contacts go to exact_kernel (path replay)
						bool cert2 = false;
						if (lane == from && !p.always_replay) {
							if (mu >= cert) {
								atomicOr(p.verdict_bits + (ps >> 5), 1u << (ps & 31u));
								verdict = TRI_SEPARATED;  // settled
								cert2 = true;
							}
						}
						if (__ballot_sync(FULL, cert2)) hm |= 1u << from;
					}
				}
			}
#endif
			// contacts whose path certificate does not hold (and, with MPTCC_NP3_DEFER_ALL, every pair the filter could
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
						p.xq[at] = make_uint4(e.z, e.x - p.cand_base[e.z].y, e.y, verdict == TRI_INTERSECT ? 1u : 0u);  // (exact_kernel counts triangles per model)
					else
						atomicOr(p.overflow, 1u);  // the fused fall-back recomputes the launch
				}
			}
			// settled pose-sets: lanes working on them drop their work; remember them for prefetched candidates and
			// queued leaf pairs
			while (hm) {
				const int from = __ffs(hm) - 1;
				hm &= hm - 1;
				const uint32_t hp = __shfl_sync(FULL, e.w, from);
				if (pose == hp) {
					has = false;
					top = 0u;
				}
				if (lane == 0) hits[hit_w & 7u] = hp;
				++hit_w;
			}
			__syncwarp();
			continue;
		}

#ifdef MPTCC_NP3_TRACE
		if (exhausted) fin_steps += (1ull << 32) + (unsigned)__popc(am);
		else run_steps += (1ull << 32) + (unsigned)__popc(am);
#endif
		// ---- node step: both children of the descended tree against the stationary node of the other ----------------
		bool leaf0 = false, leaf1 = false;
		uint2 lf0 = make_uint2(0u, 0u), lf1 = make_uint2(0u, 0u);
		unsigned n_tests = 0;
		if (has) {
			const uint2 pr = cur;
			has = false;
			// plan = (x, y | flags): x = first child of the sibling pair in the tree that is descended, y = the stationary node
			// of the other tree, kPlan3Env: the descended tree is the environment's
			const bool single = (pr.y & kPlan3Single) != 0u;
			const bool env_side = (pr.y & kPlan3Env) != 0u;
			const uint32_t xi = pr.x, yi = pr.y & kPlan3Idx;
			const float4 *rs = NP3_SLOT(lane, rd);
			const float4 r0 = rs[0], r1 = rs[1], r2 = rs[2], r3 = rs[3];  // R (9), Tc (3), Tinv (3), S
			// Loads only: nothing below consumes a loaded value before every load of the step has been issued, so the
			// two kinds of step (children = environment nodes / agent nodes) do not wait for memory one after the other.
			const unsigned second = single ? 0u : 1u;  // (the root pair is one test: its "second child" repeats the first)
			Node S;
			Node2 C2;  // both children, side by side (np_sat.cuh)
			if (TOP) {
				Node C0, C1;
				const float4 *ep = p.env.node32 + 4ull * ((unsigned long long)nb1 + (env_side ? xi : yi));
				const uint32_t ia = env_side ? yi : xi;
				if (env_side) {
					C0 = node_global(ep);
					C1 = node_global(ep + 4u * second);
					if (ia < ktop) S = node_shared_swz(atop, ia);
					else S = node_global(nodes2 + 4ull * ia);
				} else {
					S = node_global(ep);
					if (ia + 1u < ktop) {
						C0 = node_shared_swz(atop, ia);
						C1 = node_shared_swz(atop, ia + 1u);
					} else {
						C0 = node_global(nodes2 + 4ull * ia);
						C1 = node_global(nodes2 + 4ull * ia + 4);
					}
				}
				C2 = pack_nodes(C0, C1);
			} else if (Q) {
				// 32-byte quantised nodes (device_layout.hpp, nodeq): one 256-bit load per node
				const uint4 *eq = p.env.nodeq + 2ull * nb1, *aq = p.agent.nodeq + 2ull * m2.node_base;
				const uint4 *cp = (env_side ? eq : aq) + 2ull * xi, *op = (env_side ? aq : eq) + 2ull * yi;
				uint4 s0, s1, a0, a1, b0, b1;
				ldg256u(op, s0, s1);
				ldg256u(cp, a0, a1);
				ldg256u(cp + 2u * second, b0, b1);
				S = node_decode(s0, s1);
				C2 = node2_decode(a0, a1, b0, b1);
			} else {
				const float4 *e4 = p.env.node32 + 4ull * nb1;
				const float4 *cp = (env_side ? e4 : nodes2) + 4ull * xi, *op = (env_side ? nodes2 : e4) + 4ull * yi;
				S = node_global(op);
				const Node C0 = node_global(cp), C1 = node_global(cp + 4u * second);
				C2 = pack_nodes(C0, C1);
			}
			// (Ry, Ty): the relative transform when the children are environment nodes, its inverse otherwise
			const StepFrame f = step_frame(r0, r1, r2, r3, env_side, S);
			const int linkS = __float_as_int(S.v3.w);
			const float sS = S.v3.x * S.v3.x + S.v3.y * S.v3.y + S.v3.z * S.v3.z;
			float2 d2p;
			const unsigned djm = pair_disjoint<Q>(f, C2, S.v3, d2p);
			const float2 sC2 = MPTCC_FMA2(C2.e[2], C2.e[2], MPTCC_FMA2(C2.e[1], C2.e[1], MPTCC_MUL2(C2.e[0], C2.e[0])));
			n_tests = single ? 1u : 2u;
			// What the two tests lead to, in straight-line code for both children (as branches this block ran at 9-15
			// active lanes and drew 13 % of the stall samples).  Child k overlaps S iff its bit of djm is clear; two leaves
			// make a triangle pair (environment primitive, agent primitive; primitive = -(link + 1) = ~link); otherwise
			// FCL's firstOverSecond -- descend the environment node iff the agent node is a leaf, or the environment node
			// is not and is the larger one -- stated for "descend the child's tree again" (desc) against "now descend S".
			const bool lS = linkS < 0;
			const uint32_t same = yi | (env_side ? kPlan3Env : 0u), flip = env_side ? 0u : kPlan3Env;
			const bool hit0 = (djm & 1u) == 0u, hit1 = !single && (djm & 2u) == 0u;
			const bool lC0 = C2.link0 < 0, lC1 = C2.link1 < 0;
			const uint32_t pS = ~(uint32_t)linkS, pC0 = ~(uint32_t)C2.link0, pC1 = ~(uint32_t)C2.link1;
			leaf0 = hit0 && lC0 && lS;
			leaf1 = hit1 && lC1 && lS;
			lf0 = make_uint2(env_side ? pC0 : pS, env_side ? pS : pC0);
			lf1 = make_uint2(env_side ? pC1 : pS, env_side ? pS : pC1);
			const bool ok0 = hit0 && !(lC0 && lS), ok1 = hit1 && !(lC1 && lS);
			// (bitwise on purpose: as && / || the second line came back as a branch at 14 active lanes)
			const bool big0 = env_side ? sC2.x > sS : !(sS > sC2.x), big1 = env_side ? sC2.y > sS : !(sS > sC2.y);
			const bool desc0 = !lC0 & (lS | big0), desc1 = !lC1 & (lS | big1);
			const uint2 plan0 = make_uint2(desc0 ? (uint32_t)C2.link0 : (uint32_t)linkS, desc0 ? same : (xi | flip));
			const uint2 plan1 = make_uint2(desc1 ? (uint32_t)C2.link1 : (uint32_t)linkS, desc1 ? same : ((xi + 1u) | flip));
			// both children overlap: the nearer one first, the other one is stacked
			const bool both = ok0 && ok1;
			const bool first1 = both ? (p.near_first && d2p.y < d2p.x) : ok1;
			cur = first1 ? plan1 : plan0;
			has = ok0 || ok1;
			if (both) NP3_PUSH(first1 ? plan0 : plan1)
		}
		if (COUNT) {
			unsigned t = n_tests;
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
			ctr[CTR_BV] += t;
		}
		// (queue entries carry the triangle's index in the environment pack: tri_base of the candidate's model is word 17
		// of its record slot)
		const unsigned tb = __float_as_uint(NP3_SLOT(lane, rd)[4].y);  // (unconditional: one conflict-free LDS, no branch)
		{
			const unsigned lm0 = __ballot_sync(FULL, leaf0);
			if (leaf0) leafq[lq + __popc(lm0 & lt)] = make_uint4(tb + lf0.x, lf0.y, ci, pose);
			lq += __popc(lm0);
		}
		const unsigned lm = __ballot_sync(FULL, leaf1);
		if (leaf1) leafq[lq + __popc(lm & lt)] = make_uint4(tb + lf1.x, lf1.y, ci, pose);
		lq += __popc(lm);
		__syncwarp();
	}
	if (COUNT && lane == 0) {
#pragma unroll
		for (int i = 0; i < CTR_N; ++i)
			if (ctr[i]) atomicAdd(p.counters + i, ctr[i]);
	}
#ifdef MPTCC_NP3_TRACE
	if (lane == 0) {
		tr[2] = gtime();
		tr[3] = fin_steps;
		np3_trace[4ull * 8192 + ((blockIdx.x * (unsigned)kNp3Warps + (unsigned)warp) & 8191u)] = run_steps;
	}
#endif
}

}  // namespace

unsigned np3_agent_top_max();

// Agent nodes the kernel can stage beside its per-warp state (whole sibling pairs: an odd count).
unsigned np3_agent_top_max() {
	const unsigned budget = (227u * 1024u) / MPTCC_NP3_CTAS - 1024u * (MPTCC_NP3_CTAS - 1) - kNp3TopOffset - (unsigned)kNp3Warps * kNp3WarpBytes - 256u;
	unsigned n = budget / 64u;
	if (n % 2u == 0u) --n;
	return n;
}

int launch_narrowphase3(const KernelParams &p, int sm_count, bool count, void *stream) {
	const bool top = p.agent_top != 0u && !p.qnode;
	const size_t smem = kNp3TopOffset + (((size_t)(top ? p.agent_top : 0u) * 64u + 127u) & ~(size_t)127u) + (size_t)kNp3Warps * kNp3WarpBytes;
	// (the opt-in is made once for the largest block the kernel can ask for: prepare_kernel caches per size, and a
	// later opt-in to a smaller size would lower the kernel's limit again)
	const size_t smem_max = kNp3TopOffset + (((size_t)np3_agent_top_max() * 64u + 127u) & ~(size_t)127u) + (size_t)kNp3Warps * kNp3WarpBytes;
	if (smem > smem_max) return (int)cudaErrorInvalidValue;
	void (*k)(KernelParams) = p.qnode ? (count ? narrowphase3_kernel<true, false, true> : narrowphase3_kernel<false, false, true>)
	                          : top   ? (count ? narrowphase3_kernel<true, true, false> : narrowphase3_kernel<false, true, false>)
	                                  : (count ? narrowphase3_kernel<true, false, false> : narrowphase3_kernel<false, false, false>);
	int per_sm = 0;
	const cudaError_t e = prepare_kernel(reinterpret_cast<const void *>(k), kNp3Warps * 32, top ? smem_max : smem, &per_sm);
	if (e != cudaSuccess) return (int)e;
	if (per_sm < 1) per_sm = 1;
	k<<<sm_count * per_sm, kNp3Warps * 32, smem, static_cast<cudaStream_t>(stream)>>>(p);
#ifdef MPTCC_NP3_TRACE
	{
		static int calls = 0;
		if (++calls % 8 == 0) {
			const int nw = std::min(8192, sm_count * per_sm * kNp3Warps);
			std::vector<unsigned long long> h(4 * (size_t)nw), h2(nw);
			cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
			cudaMemcpyFromSymbol(h.data(), np3_trace, h.size() * 8);
			cudaMemcpyFromSymbol(h2.data(), np3_trace, h2.size() * 8, 4ull * 8192 * 8);
			unsigned long long t0 = ~0ull, fi = 0, fl = 0, ri = 0, rl = 0;
			for (int w = 0; w < nw; ++w) {
				t0 = std::min(t0, h[4 * w]);
				fi += h[4 * w + 3] >> 32;
				fl += h[4 * w + 3] & 0xffffffffull;
				ri += h2[w] >> 32;
				rl += h2[w] & 0xffffffffull;
			}
			fprintf(stderr, "np3 trace: node steps before the list ran out %.1f per warp at %.1f lanes, after %.1f per warp at %.1f lanes\n", (double)ri / nw,
			        (double)rl / (double)(ri ? ri : 1), (double)fi / nw, (double)fl / (double)(fi ? fi : 1));
			std::vector<double> ex, en;
			for (int w = 0; w < nw; ++w) {
				ex.push_back(h[4 * w + 1] ? (h[4 * w + 1] - t0) * 1e-3 : -1.0);
				en.push_back((h[4 * w + 2] - t0) * 1e-3);
			}
			std::sort(ex.begin(), ex.end());
			std::sort(en.begin(), en.end());
			auto pc = [&](const std::vector<double> &v, double q) { return v[(size_t)(q * (v.size() - 1))]; };
			fprintf(stderr, "np3 trace (us after the first warp started), %d warps: list ran out min %.1f 50%% %.1f max %.1f | exit min %.1f 10%% %.1f 50%% %.1f 90%% %.1f 99%% %.1f max %.1f\n", nw,
			        pc(ex, 0), pc(ex, .5), pc(ex, 1), pc(en, 0), pc(en, .1), pc(en, .5), pc(en, .9), pc(en, .99), pc(en, 1));
		}
	}
#endif
	return (int)cudaGetLastError();
}

}  // namespace mptcc
