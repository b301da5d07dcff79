// collide_kernel.cu -- the fused kernel of libmptcc: broadphase + narrowphase of a chunk of pose-sets in
// one warp (see collide_device.cuh for the algorithm).  Used for small batches, multi-link agents and as the
// overflow fall-back of the split pipeline (split_kernels.cu).
#include "collide_device.cuh"
#include "launch_cache.hpp"

namespace mptcc {
namespace {

template <bool COUNT>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) collide_kernel(const __grid_constant__ KernelParams p) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	unsigned char *base = smem_raw + (size_t)warp * warp_smem_bytes();
	WarpMem wm;
	wm.leafq = reinterpret_cast<uint2 *>(base);
	base += kLeafQueueCap * 8;
	wm.cand = reinterpret_cast<float4 *>(base);
	base += kCandCap * 64;
	wm.box = reinterpret_cast<double *>(base);
	base += 6 * 32 * 8;
	wm.rel = reinterpret_cast<double *>(base);
	base += 16 * 8;
	wm.cand_ids = reinterpret_cast<uint32_t *>(base);
	base += kCandCap * 4;
	wm.clq = reinterpret_cast<uint32_t *>(base);
	base += kClusterQueueCap * 4;
	wm.donor = base;

	if (p.run_if != nullptr && *p.run_if == 0u) return;  // overflow fall-back of the split pipeline: nothing to do
	const unsigned ppw = 1u << p.ppw_log2;
	const unsigned long long n_chunks = (p.n_pose_sets + ppw - 1) >> p.ppw_log2;
	const unsigned rec_len = pose_rec_len(p.pose_format);
	const bool aligned32 = (reinterpret_cast<unsigned long long>(p.poses) & 31ull) == 0ull;
	const unsigned lt = (1u << lane) - 1u;
	unsigned long long ctr[CTR_N];
#pragma unroll
	for (int i = 0; i < CTR_N; ++i) ctr[i] = 0;

	bool first_turn = true;
	while (true) {
		unsigned chunk = 0;
		if (p.work_counter != nullptr) {
			if (lane == 0) chunk = atomicAdd(p.work_counter, 1u);
			chunk = __shfl_sync(FULL, chunk, 0);
		} else {
			// planner-sized launches: one chunk per warp, the grid is sized to fit -- no work counter, hence no memset
			// in front of the launch (a CUDA call less on a path that is all launch latency)
			if (!first_turn) break;
			first_turn = false;
			chunk = blockIdx.x * kWarpsPerCta + warp;
		}
		if (chunk >= n_chunks) break;
		const unsigned long long pose0 = (unsigned long long)chunk << p.ppw_log2;
		const unsigned long long left = p.n_pose_sets - pose0;
		const unsigned nvalid = left < ppw ? (unsigned)left : ppw;
		const u64 validmask = nvalid >= 64u ? ~0ull : ((1ull << nvalid) - 1ull);
		const unsigned rounds = (nvalid + 31u) >> 5;  // 32 pose slots per round, up to 64 slots per chunk
		FusedTrav tv;
		tv.hitmask = 0;
		tv.validmask = validmask;
		tv.stride = (size_t)rec_len * p.n_links;
		u64 &hitmask = tv.hitmask;

		for (unsigned j = 0; j < p.n_links && hitmask != validmask; ++j) {
			const DevModel &m2 = p.agent.models[p.link_model[j]];
			const double *link_poses = p.poses + (pose0 * p.n_links + j) * rec_len;  // pose slot s: + s*rec_len*n_links
			tv.link_poses = link_poses;
			int ncand = 0, nrec = 0;  // the candidate table is shared by the rounds of this link
			for (unsigned round = 0; round < rounds && hitmask != validmask; ++round) {
				const unsigned slot = round * 32u + (unsigned)lane;
				const bool valid = slot < nvalid;
				// ---- phase A: lane-per-pose decode + world AABB (CollisionObject::computeAABB) -------
				float amin[3] = {0.f, 0.f, 0.f}, amax[3] = {0.f, 0.f, 0.f};
				if (valid) {
					double rec[12];
					load_pose_record(link_poses + (size_t)slot * rec_len * p.n_links, p.pose_format, aligned32, rec);
					pose_world_box(rec, p.pose_format, m2, wm.box + lane, amin, amax);
				}
				int qmin[3], qmax[3];
				quantize_pose_box(p, amin, amax, qmin, qmax);
				__syncwarp();
				unsigned node = valid ? 0u : p.n_bp_nodes;
				int ncl = 0;
				bool more = true;
				while (more || ncl > 0) {
					// ---- phase B: lane-per-pose walk of the threaded cluster tree (fp32, conservative) ----
					while (more && ncl <= kClusterQueueCap - 32) {
						if ((hitmask >> slot) & 1ull) node = p.n_bp_nodes;
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
					// ---- phase C1: (pose, cluster) -> exact fp64 AABB::overlap per object -> candidate ids -
					// 32 / kClusterSize queue entries per iteration: lane -> (entry, object of the cluster)
					constexpr int kPer = 32 / kClusterSize;
					int c = 0;
					for (; c < ncl && ncand <= kCandCap - 32; c += kPer) {
						const int idx = c + lane / kClusterSize;
						bool ov = false;
						unsigned src = 0, eobj = 0;
						if (idx < ncl) {
							const unsigned ent = wm.clq[idx];
							const unsigned sl = ent >> 27;  // lane slot of this round
							src = round * 32u + sl;
							eobj = (ent & 0x7ffffffu) * kClusterSize + (unsigned)(lane % kClusterSize);
							if (!((hitmask >> src) & 1ull)) ov = exact_aabb_overlap(p, wm, eobj, sl);
						}
						const unsigned om = __ballot_sync(FULL, ov);
						if (ov) wm.cand_ids[ncand + __popc(om & lt)] = (src << 26) | eobj;
						ncand += __popc(om);
					}
					if (c > ncl) c = ncl;
					// keep the unprocessed tail of the cluster queue
					if (c < ncl) {
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
					if (COUNT) ctr[CTR_CAND] += (unsigned)(ncand - nrec);
					// ---- phase C2: one lane per new candidate: fp64 relative transform -> fp32 record -------
					for (int i0 = nrec; i0 < ncand; i0 += 32) {
						const int i = i0 + lane;
						if (i < ncand) {
							const uint32_t id = wm.cand_ids[i];
							const unsigned src = id >> 26, e = id & 0x3ffffffu;
							float4 rec[4];
							double pose[12];
							load_pose_record(link_poses + (size_t)src * rec_len * p.n_links, p.pose_format, aligned32, pose);
							make_candidate_record(p, pose, p.objects[e], m2, src, rec);
#pragma unroll
							for (unsigned k = 0; k < 4; ++k) wm.cand[rec_at((unsigned)i, k)] = rec[k];
						}
					}
					nrec = ncand;
					__syncwarp();
					// ---- phase D (table full): merged narrowphase ---------------------------------------------
					if (ncand > kCandCap - 32) {
						traverse_all<COUNT>(p, wm, m2, ncand, tv, lane, ctr);
						ncand = 0;
						nrec = 0;
						if (hitmask == validmask) break;
					}
				}
			}
			// ---- phase D: the candidates left after the last round of this link ---------------------------
			if (ncand > 0 && hitmask != validmask) traverse_all<COUNT>(p, wm, m2, ncand, tv, lane, ctr);
		}
		// ---- verdict bits ----------------------------------------------------------------------
		if (lane == 0) {
			const unsigned long long word = pose0 >> 5;
			if (ppw >= 32u) {
				p.verdict_bits[word] = (uint32_t)hitmask;
				if (nvalid > 32u) p.verdict_bits[word + 1] = (uint32_t)(hitmask >> 32);
			} else if (hitmask) {
				atomicOr(p.verdict_bits + word, (uint32_t)hitmask << (unsigned)(pose0 & 31ull));
			}
		}
	}
	if (COUNT && lane == 0) {
#pragma unroll
		for (int i = 0; i < CTR_N; ++i)
			if (ctr[i]) atomicAdd(p.counters + i, ctr[i]);
	}
}

}  // namespace

int collide_max_ctas_per_sm(bool count) {
	int n = 0;
	const size_t smem = (size_t)kWarpsPerCta * warp_smem_bytes();
	const void *k = count ? reinterpret_cast<const void *>(collide_kernel<true>) : reinterpret_cast<const void *>(collide_kernel<false>);
	if (prepare_kernel(k, kWarpsPerCta * 32, smem, &n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

int launch_collide(const KernelParams &p, int grid_ctas, bool count, void *stream) {
	const size_t smem = (size_t)kWarpsPerCta * warp_smem_bytes();
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const void *k = count ? reinterpret_cast<const void *>(collide_kernel<true>) : reinterpret_cast<const void *>(collide_kernel<false>);
	const cudaError_t e = prepare_kernel(k, kWarpsPerCta * 32, smem, nullptr);
	if (e != cudaSuccess) return (int)e;
	if (count)
		collide_kernel<true><<<grid_ctas, kWarpsPerCta * 32, smem, s>>>(p);
	else
		collide_kernel<false><<<grid_ctas, kWarpsPerCta * 32, smem, s>>>(p);
	return (int)cudaGetLastError();
}

}  // namespace mptcc
