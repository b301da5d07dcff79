// split_kernels.cu -- the two-kernel pipeline of libmptcc for large batches of single-link pose-sets.
//
//   broadphase_kernel  : phases A-C2 of collide_device.cuh for chunks of 32 pose-sets.  It has no traversal
//                        state, so it runs at a much higher occupancy than the fused kernel and hides the
//                        latency of its dependent loads (pose records, tree walk, object boxes) with warps.
//                        Output: a global list of candidates (pose-set, environment object, fp32 record, fp64
//                        relative transform), the candidates of one pose-set adjacent.
//   narrowphase_kernel : persistent warps take batches of kCandCap candidates from that list, whatever
//                        pose-sets they belong to, and run the depth-first lane traversal (traverse_all) on
//                        them: no chunk boundaries, always a full table, perfect load balance.  Hits are
//                        published with atomicOr into the (zeroed) verdict words.
//
// If the list overflows its capacity the broadphase sets *overflow and the host-enqueued fused kernel, which
// is launched right after with run_if = overflow, recomputes the launch (verdict bits only ever get set for
// colliding pose-sets, so OR-ing both results is exact).
#include "collide_device.cuh"
#include "launch_cache.hpp"

namespace mptcc {
namespace {

constexpr int kBpWarpsPerCta = 8;

// Candidates a narrowphase warp traverses together.  Long batches keep the lanes busy longer before the
// end-of-batch stealing phase (+3.8 % on a 2^20-pose launch at 128 instead of 96) but a short candidate list
// then has fewer batches than warps, so the kernel is built for two lengths and the host picks per launch.
#ifndef MPTCC_NP_BATCH
#define MPTCC_NP_BATCH MPTCC_CAND_CAP
#endif
#ifndef MPTCC_NP_BATCH_LONG
#define MPTCC_NP_BATCH_LONG 128
#endif
constexpr int kNpBatch = MPTCC_NP_BATCH, kNpBatchLong = MPTCC_NP_BATCH_LONG;

// per-warp shared memory of the broadphase kernel: pose boxes + cluster queue + candidate ids
__host__ __device__ constexpr unsigned bp_warp_smem_bytes() { return 6u * 32u * 8u + kClusterQueueCap * 4u + kCandCap * 4u + 32u * 4u; }
// ... of the grid broadphase kernel: + the decoded poses of the chunk
__host__ __device__ constexpr unsigned bp_grid_warp_smem_bytes() { return bp_warp_smem_bytes() + 12u * 32u * 8u; }

// per-warp shared memory of the narrowphase kernel
__host__ __device__ constexpr unsigned np_warp_smem_bytes(unsigned batch) {
	return kLeafQueueCap * 8u + batch * 64u + 16u * 8u + batch * 4u + batch * 4u + batch + 32u;
}

// Phase C2 + append: the `ncand` candidates (pose slot << 26 | object) in wm.cand_ids are grouped by pose slot, their
// records are computed (fp64 relative transform -> fp32 record) and appended to the global candidate list.
// SORTED: the candidates arrive grouped by pose slot already (grid broadphase) and `rt` holds the decoded poses of the
// chunk (R 9, T 3, component-major with stride 32): no regrouping, no second decode.
template <bool NP2, bool SORTED = false>
__device__ __forceinline__ void flush_candidates(const KernelParams &p, const WarpMem &wm, unsigned *pfx, int ncand,
                                                 const double *chunk_poses, unsigned rec_len, bool aligned32,
                                                 unsigned long long pose0, const DevModel &m2, int lane,
                                                 const double *rt = nullptr) {
	// (the reservation first: its round trip to L2 overlaps the counting below)
	unsigned gbase = 0;
	if (lane == 0) gbase = atomicAdd(p.cand_total, (unsigned)ncand);
	if (!SORTED) {
		// lane = pose slot: count its candidates; exclusive prefix over the slots -> pfx[]
		int mine = 0;
		for (int i = 0; i < ncand; ++i) mine += (wm.cand_ids[i] >> 26) == (unsigned)lane;
		int pre = mine;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const int v = __shfl_up_sync(FULL, pre, o);
			if (lane >= o) pre += v;
		}
		pfx[lane] = (unsigned)(pre - mine);
	}
	gbase = __shfl_sync(FULL, gbase, 0);
	__syncwarp();
	if ((unsigned long long)gbase + (unsigned)ncand > (unsigned long long)p.cand_cap) {
		if (lane == 0) atomicOr(p.overflow, 1u);  // the fused fall-back recomputes this launch
		return;
	}
	for (int i = lane; i < ncand; i += 32) {
		const uint32_t id = wm.cand_ids[i];
		const unsigned sl = id >> 26, e = id & 0x3ffffffu;
		unsigned dst = gbase + (unsigned)i;
		if (!SORTED) {
			int rank = 0;  // candidates of the same pose slot that precede i
			for (int k = 0; k < i; ++k) rank += (wm.cand_ids[k] >> 26) == sl;
			dst = gbase + pfx[sl] + (unsigned)rank;
		}
		float4 rec[4];
		double pose[12], rel[12];
		if (!(SORTED && NP2)) load_pose_record(chunk_poses + (size_t)sl * rec_len, p.pose_format, aligned32, pose);
		float4 *out = p.cand_rec + 4ull * dst;
		if (NP2) {
			uint2 mbase;
			float cert;
			bool rigid;
			if (SORTED) {
				double R2[9], T2[3];
#pragma unroll
				for (int k = 0; k < 9; ++k) R2[k] = rt[k * 32 + sl];
#pragma unroll
				for (int k = 0; k < 3; ++k) T2[k] = rt[(9 + k) * 32 + sl];
				rigid = make_candidate_record2_rt(p, R2, T2, p.objects[e], m2, rec, rel, mbase, &cert);
			} else {
				rigid = make_candidate_record2(p, pose, p.objects[e], m2, rec, rel, mbase, &cert);
			}
			if (!rigid)
				atomicOr(p.overflow, 2u);  // non-rigid pose: the fused fall-back recomputes the launch
			p.cand_base[dst] = mbase;
			p.cand_cert[dst] = cert;
		} else {
			make_candidate_record(p, pose, p.objects[e], m2, 0u, rec, rel);
		}
		// FCL's fp64 relative transform goes along for the exact pass.  (Recomputing it there from the pose
		// record instead was measured: the broadphase kernel is not faster without these 96 bytes per
		// candidate, and the narrowphase kernel is 3 % slower.)
		double *orel = p.cand_rel + 12ull * dst;
		stg256(orel, rel[0], rel[1], rel[2], rel[3]);
		stg256(orel + 4, rel[4], rel[5], rel[6], rel[7]);
		stg256(orel + 8, rel[8], rel[9], rel[10], rel[11]);
#ifdef MPTCC_LIST_EF
		stg256_evict_first(out, rec[0], rec[1]);
		stg256_evict_first(out + 2, rec[2], rec[3]);
#else
		stg256(out, rec[0], rec[1]);
		stg256(out + 2, rec[2], rec[3]);
#endif
		p.cand_obj[dst] = e;
		p.cand_pose[dst] = (uint32_t)(pose0 + sl);
	}
}

template <bool COUNT, bool NP2>
__global__ void __launch_bounds__(kBpWarpsPerCta * 32, 3) broadphase_kernel(const __grid_constant__ KernelParams p) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	unsigned char *base = smem_raw + (size_t)warp * bp_warp_smem_bytes();
	WarpMem wm{};
	wm.box = reinterpret_cast<double *>(base);
	base += 6 * 32 * 8;
	wm.clq = reinterpret_cast<uint32_t *>(base);
	base += kClusterQueueCap * 4;
	wm.cand_ids = reinterpret_cast<uint32_t *>(base);
	base += kCandCap * 4;
	unsigned *pfx = reinterpret_cast<unsigned *>(base);  // [32] first list position of each pose slot's run

	const unsigned long long n_chunks = (p.n_pose_sets + 31ull) >> 5;
	const unsigned rec_len = pose_rec_len(p.pose_format);
	const bool aligned32 = (reinterpret_cast<unsigned long long>(p.poses) & 31ull) == 0ull;
	const unsigned lt = (1u << lane) - 1u;
	const DevModel &m2 = p.agent.models[p.link_model[0]];
	unsigned long long n_cand_total = 0;

	while (true) {
		unsigned chunk = 0;
		if (lane == 0) chunk = atomicAdd(p.work_counter, 1u);
		chunk = __shfl_sync(FULL, chunk, 0);
		if (chunk >= n_chunks) break;
		const unsigned long long pose0 = (unsigned long long)chunk << 5;
		const unsigned long long left = p.n_pose_sets - pose0;
		const unsigned nvalid = left < 32ull ? (unsigned)left : 32u;
		const bool valid = (unsigned)lane < nvalid;
		const double *chunk_poses = p.poses + pose0 * rec_len;

		// ---- phase A ------------------------------------------------------------------------------------
		float amin[3] = {0.f, 0.f, 0.f}, amax[3] = {0.f, 0.f, 0.f};
		if (valid) {
			double rec[12];
			load_pose_record(chunk_poses + (size_t)lane * rec_len, p.pose_format, aligned32, rec);
			pose_world_box(rec, p.pose_format, m2, wm.box + lane, amin, amax);
		}
		int qmin[3], qmax[3];
		quantize_pose_box(p, amin, amax, qmin, qmax);
		__syncwarp();
		unsigned node = valid ? 0u : p.n_bp_nodes;
		int ncl = 0, ncand = 0;
		bool more = true;
		while (more || ncl > 0 || ncand > 0) {
			// ---- phase B ----------------------------------------------------------------------------------
			while (more && ncl <= kClusterQueueCap - 32) {
				const bool act = node < p.n_bp_nodes;
				if (!__any_sync(FULL, act)) {
					more = false;
					break;
				}
				unsigned cl = 0;
				const bool leaf_hit = act && bp_tree_step(p, node, qmin, qmax, cl);
				const unsigned m = __ballot_sync(FULL, leaf_hit);
				if (m) {
					if (leaf_hit) wm.clq[ncl + __popc(m & lt)] = ((unsigned)lane << 27) | cl;
					ncl += __popc(m);
				}
			}
			__syncwarp();
			// ---- phase C1 ---------------------------------------------------------------------------------
			constexpr int kPer = 32 / kClusterSize;
			int c = 0;
			for (; c < ncl && ncand <= kCandCap - 32; c += kPer) {
				const int idx = c + lane / kClusterSize;
				bool ov = false;
				unsigned sl = 0, eobj = 0;
				if (idx < ncl) {
					const unsigned ent = wm.clq[idx];
					sl = ent >> 27;
					eobj = (ent & 0x7ffffffu) * kClusterSize + (unsigned)(lane % kClusterSize);
					ov = exact_aabb_overlap(p, wm, eobj, sl);
				}
				const unsigned om = __ballot_sync(FULL, ov);
				if (ov) wm.cand_ids[ncand + __popc(om & lt)] = (sl << 26) | eobj;
				ncand += __popc(om);
			}
			if (c > ncl) c = ncl;
			if (c < ncl) {  // keep the unprocessed tail of the cluster queue
				const int rest = ncl - c;
				unsigned keep0 = 0, keep1 = 0;
				if (lane < rest) keep0 = wm.clq[c + lane];
				if (lane + 32 < rest) keep1 = wm.clq[c + lane + 32];
				__syncwarp();
				if (lane < rest) wm.clq[lane] = keep0;
				if (lane + 32 < rest) wm.clq[lane + 32] = keep1;
				ncl = rest;
			} else {
				ncl = 0;
			}
			__syncwarp();
			// ---- flush: group by pose slot, phase C2, append to the global candidate list --------------------
			if (ncand > kCandCap - 32 || (!more && ncl == 0)) {
				if (ncand > 0) {
					flush_candidates<NP2>(p, wm, pfx, ncand, chunk_poses, rec_len, aligned32, pose0, m2, lane);
					n_cand_total += (unsigned)ncand;
				}
				ncand = 0;
				__syncwarp();
			}
		}
	}
	if (COUNT && lane == 0 && n_cand_total) atomicAdd(p.counters + CTR_CAND, n_cand_total);
}


#ifndef MPTCC_BP_GRID_CTAS
#define MPTCC_BP_GRID_CTAS 3
#endif
// Grid broadphase (device_layout.hpp, "Broadphase grid"): the world box of a pose of a single-link agent spans at most
// two cells per axis, and every cell carries the list of the objects that touch its 2 x 2 x 2 block, with their exact
// fp64 boxes inline: one header load and one coalesced stream of 64-byte entries per pose instead of a tree walk.
// The list is a superset of FCL's candidates (the cell index is the same monotone function of a coordinate on the host
// and here); every entry is re-tested with FCL's exact fp64 AABB::overlap.
template <bool COUNT, bool NP2>
__global__ void __launch_bounds__(kBpWarpsPerCta * 32, MPTCC_BP_GRID_CTAS) broadphase_grid_kernel(const __grid_constant__ KernelParams p) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	unsigned char *base = smem_raw + (size_t)warp * bp_grid_warp_smem_bytes();
	WarpMem wm{};
	wm.box = reinterpret_cast<double *>(base);
	base += 6 * 32 * 8;
	unsigned *excl = reinterpret_cast<unsigned *>(base);  // [32] first flat index of each pose slot's entries (cluster queue space)
	unsigned *first = excl + 32;                           // [32] first list entry of each pose slot
	base += kClusterQueueCap * 4;
	wm.cand_ids = reinterpret_cast<uint32_t *>(base);
	base += kCandCap * 4;
	unsigned *pfx = reinterpret_cast<unsigned *>(base);
	base += 32 * 4;
	double *rt = reinterpret_cast<double *>(base);  // [12][32] decoded pose (R, T) of every lane's pose

	const unsigned long long n_chunks = (p.n_pose_sets + 31ull) >> 5;
	const unsigned rec_len = pose_rec_len(p.pose_format);
	const bool aligned32 = (reinterpret_cast<unsigned long long>(p.poses) & 31ull) == 0ull;
	const unsigned lt = (1u << lane) - 1u;
	const DevModel &m2 = p.agent.models[p.link_model[0]];
	unsigned long long n_cand_total = 0;

	while (true) {
		unsigned chunk = 0;
		if (lane == 0) chunk = atomicAdd(p.work_counter, 1u);
		chunk = __shfl_sync(FULL, chunk, 0);
		if (chunk >= n_chunks) break;
		const unsigned long long pose0 = (unsigned long long)chunk << 5;
		const unsigned long long left = p.n_pose_sets - pose0;
		const unsigned nvalid = left < 32ull ? (unsigned)left : 32u;
		const bool valid = (unsigned)lane < nvalid;
		const double *chunk_poses = p.poses + pose0 * rec_len;

		// ---- phase A + cell look-up (lane = pose) -------------------------------------------------------------------
		unsigned start = 0, count = 0;
		bool bad = false;
		if (valid) {
			double rec[12];
			float amin[3], amax[3];
			load_pose_record(chunk_poses + (size_t)lane * rec_len, p.pose_format, aligned32, rec);
			pose_world_box(rec, p.pose_format, m2, wm.box + lane, amin, amax, rt + lane);
			int c[3];
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				const double lo = wm.box[k * 32 + lane], hi = wm.box[(3 + k) * 32 + lane];
				const int a = grid_cell(lo, p.grid_origin[k], p.grid_inv[k], p.grid_n[k]);
				const int b = grid_cell(hi, p.grid_origin[k], p.grid_inv[k], p.grid_n[k]);
				// a box that is not finite, or spans more than two cells (coordinates too large for the fp32 cell function):
				// the launch goes to the fused kernel, which walks the tree with FCL's semantics for such boxes
				bad = bad || !(fabs(lo) <= 1.7e308) || !(fabs(hi) <= 1.7e308) || b - a > 1 || b < a;
				c[k] = a;
			}
			if (!bad) {
				const uint2 hdr = __ldg(p.grid_cells + ((size_t)c[0] * p.grid_n[1] + c[1]) * p.grid_n[2] + c[2]);
				start = hdr.x;
				count = hdr.y;
			}
		}
		if (__any_sync(FULL, bad)) {
			if (lane == 0) atomicOr(p.overflow, 2u);
			break;  // (every warp that sees the flag stops early; the fall-back recomputes everything)
		}
		unsigned pre = count;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned v = __shfl_up_sync(FULL, pre, o);
			if (lane >= o) pre += v;
		}
		const unsigned total = __shfl_sync(FULL, pre, 31);
		excl[lane] = pre - count;
		first[lane] = start;
		__syncwarp();
		// ---- exact fp64 AABB::overlap of every (pose, list entry), 32 at a time -----------------------------------------
		int ncand = 0;
		for (unsigned t0 = 0; t0 < total; t0 += 32u) {
			const unsigned t = t0 + (unsigned)lane;
			bool ov = false;
			unsigned sl = 0, eobj = 0;
			if (t < total) {
				// slot = the last one whose first flat index is <= t
#pragma unroll
				for (unsigned s = 16u; s > 0u; s >>= 1)
					if (excl[sl + s] <= t) sl += s;
				const GridEntry *ge = p.grid_entries + (first[sl] + (t - excl[sl]));
				double e[8];
				asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(e[0]), "=d"(e[1]), "=d"(e[2]), "=d"(e[3]) : "l"(ge));
				asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(e[4]), "=d"(e[5]), "=d"(e[6]), "=d"(e[7]) : "l"(reinterpret_cast<const double *>(ge) + 4));
				eobj = (unsigned)__double_as_longlong(e[6]);
				const bool out = (e[0] > wm.box[3 * 32 + sl]) | (e[1] > wm.box[4 * 32 + sl]) | (e[2] > wm.box[5 * 32 + sl]) |
				                 (e[3] < wm.box[0 * 32 + sl]) | (e[4] < wm.box[1 * 32 + sl]) | (e[5] < wm.box[2 * 32 + sl]);
				ov = !out;
			}
			const unsigned om = __ballot_sync(FULL, ov);
			if (ov) wm.cand_ids[ncand + __popc(om & lt)] = (sl << 26) | eobj;
			ncand += __popc(om);
			if (ncand > kCandCap - 32) {
				__syncwarp();
				flush_candidates<NP2, true>(p, wm, pfx, ncand, chunk_poses, rec_len, aligned32, pose0, m2, lane, rt);
				n_cand_total += (unsigned)ncand;
				ncand = 0;
				__syncwarp();
			}
		}
		__syncwarp();
		if (ncand > 0) {
			flush_candidates<NP2, true>(p, wm, pfx, ncand, chunk_poses, rec_len, aligned32, pose0, m2, lane, rt);
			n_cand_total += (unsigned)ncand;
		}
		__syncwarp();
	}
	if (COUNT && lane == 0 && n_cand_total) atomicAdd(p.counters + CTR_CAND, n_cand_total);
}


#ifndef MPTCC_NP_WARPS
#define MPTCC_NP_WARPS 8
#endif
#ifndef MPTCC_NP_CTAS
#define MPTCC_NP_CTAS 2
#endif
constexpr int kNpWarpsPerCta = MPTCC_NP_WARPS;

#ifdef MPTCC_NP_MAXREG
#define NP_BOUNDS __maxnreg__(MPTCC_NP_MAXREG)
#else
#define NP_BOUNDS __launch_bounds__(kNpWarpsPerCta * 32, MPTCC_NP_CTAS)
#endif

template <bool COUNT, int BATCH>
__global__ void NP_BOUNDS narrowphase_kernel(const __grid_constant__ KernelParams p) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	unsigned char *base = smem_raw + (size_t)warp * np_warp_smem_bytes(BATCH);
	WarpMem wm{};
	wm.leafq = reinterpret_cast<uint2 *>(base);
	base += kLeafQueueCap * 8;
	wm.cand = reinterpret_cast<float4 *>(base);
	base += BATCH * 64;
	wm.rel = reinterpret_cast<double *>(base);
	base += 16 * 8;
	wm.cand_ids = reinterpret_cast<uint32_t *>(base);
	base += BATCH * 4;
	wm.cand_pose = reinterpret_cast<uint32_t *>(base);
	base += BATCH * 4;
	wm.dead = base;
	base += BATCH;
	wm.donor = base;

	// An overflowed list has holes (a flush that did not fit wrote nothing): leave the whole launch to the
	// fused fall-back kernel.
	if (*p.overflow != 0u) return;
	const unsigned total = *p.cand_total;
	const DevModel &m2 = p.agent.models[p.link_model[0]];
	SplitTrav tv;
	tv.poses = p.poses;
	tv.rec_len = pose_rec_len(p.pose_format);
	tv.verdict_bits = p.verdict_bits;
	unsigned long long ctr[CTR_N];
#pragma unroll
	for (int i = 0; i < CTR_N; ++i) ctr[i] = 0;

	while (true) {
		unsigned first = 0;
		if (lane == 0) first = atomicAdd(p.work_counter2, (unsigned)BATCH);
		first = __shfl_sync(FULL, first, 0);
		if (first >= total) break;
		const int n = (int)(total - first < (unsigned)BATCH ? total - first : (unsigned)BATCH);
		tv.batch_first = first;
		// ---- load the batch: records (coalesced float4 copy), objects, pose-set indices -------------------
		const float4 *src_rec = p.cand_rec + 4ull * first;
		for (int i = lane; i < 4 * n; i += 32) wm.cand[rec_at((unsigned)i >> 2, (unsigned)i & 3u)] = src_rec[i];
		for (int i = lane; i < n; i += 32) {
			wm.cand_ids[i] = p.cand_obj[first + i];
			wm.cand_pose[i] = p.cand_pose[first + i];
		}
		__syncwarp();
		// ---- pose slots of the batch: candidates of one pose-set are adjacent; the slot of a run is the
		//      index of its first candidate.  dead[slot] starts from the verdict bit (another batch may have
		//      settled the pose-set already). -----------------------------------------------------------------
		unsigned carry = 0;
		for (int g = 0; g < (n + 31) / 32; ++g) {
			const int i = 32 * g + lane;
			const bool valid = i < n;
			const uint32_t pose = valid ? wm.cand_pose[i] : 0u;
			const bool start = valid && (i == 0 || wm.cand_pose[i - 1] != pose);
			const unsigned sm = __ballot_sync(FULL, start);
			const unsigned upto = sm & (0xffffffffu >> (31 - lane));
			const unsigned slot = upto ? (unsigned)(32 * g + 31 - __clz((int)upto)) : carry;
			if (valid) wm.cand[rec_at((unsigned)i, 3u)].y = __uint_as_float(slot);
			if (start) wm.dead[i] = (unsigned char)((p.verdict_bits[pose >> 5] >> (pose & 31u)) & 1u);
			if (sm) carry = (unsigned)(32 * g + 31 - __clz((int)sm));
		}
		__syncwarp();
		traverse_all<COUNT>(p, wm, m2, n, tv, lane, ctr);
		__syncwarp();
	}
	if (COUNT && lane == 0) {
#pragma unroll
		for (int i = 0; i < CTR_N; ++i)
			if (ctr[i]) atomicAdd(p.counters + i, ctr[i]);
	}
}

template <class K>
int launch_persistent(K kernel, const KernelParams &p, int sm_count, int threads, size_t smem, void *stream) {
	int per_sm = 0;
	const cudaError_t e = prepare_kernel(reinterpret_cast<const void *>(kernel), threads, smem, &per_sm);
	if (e != cudaSuccess) return (int)e;
	if (per_sm < 1) per_sm = 1;
	kernel<<<sm_count * per_sm, threads, smem, static_cast<cudaStream_t>(stream)>>>(p);
	return (int)cudaGetLastError();
}

}  // namespace

int launch_broadphase(const KernelParams &p, int sm_count, bool count, bool np2, void *stream) {
	const size_t smem = (size_t)kBpWarpsPerCta * bp_warp_smem_bytes();
	if (p.grid_cells != nullptr) {
		const size_t smem = (size_t)kBpWarpsPerCta * bp_grid_warp_smem_bytes();
		if (np2)
			return count ? launch_persistent(broadphase_grid_kernel<true, true>, p, sm_count, kBpWarpsPerCta * 32, smem, stream)
			             : launch_persistent(broadphase_grid_kernel<false, true>, p, sm_count, kBpWarpsPerCta * 32, smem, stream);
		return count ? launch_persistent(broadphase_grid_kernel<true, false>, p, sm_count, kBpWarpsPerCta * 32, smem, stream)
		             : launch_persistent(broadphase_grid_kernel<false, false>, p, sm_count, kBpWarpsPerCta * 32, smem, stream);
	}
	if (np2)
		return count ? launch_persistent(broadphase_kernel<true, true>, p, sm_count, kBpWarpsPerCta * 32, smem, stream)
		             : launch_persistent(broadphase_kernel<false, true>, p, sm_count, kBpWarpsPerCta * 32, smem, stream);
	return count ? launch_persistent(broadphase_kernel<true, false>, p, sm_count, kBpWarpsPerCta * 32, smem, stream)
	             : launch_persistent(broadphase_kernel<false, false>, p, sm_count, kBpWarpsPerCta * 32, smem, stream);
}

int launch_narrowphase(const KernelParams &p, int sm_count, bool count, bool long_batches, void *stream) {
	const int threads = kNpWarpsPerCta * 32;
	if (long_batches) {
		const size_t smem = (size_t)kNpWarpsPerCta * np_warp_smem_bytes(kNpBatchLong);
		return count ? launch_persistent(narrowphase_kernel<true, kNpBatchLong>, p, sm_count, threads, smem, stream)
		             : launch_persistent(narrowphase_kernel<false, kNpBatchLong>, p, sm_count, threads, smem, stream);
	}
	const size_t smem = (size_t)kNpWarpsPerCta * np_warp_smem_bytes(kNpBatch);
	return count ? launch_persistent(narrowphase_kernel<true, kNpBatch>, p, sm_count, threads, smem, stream)
	             : launch_persistent(narrowphase_kernel<false, kNpBatch>, p, sm_count, threads, smem, stream);
}

}  // namespace mptcc
