// self_kernel.cu -- checkSelfCollision = true (/root/reference/utilities/meshhandler.hpp:205-219): do the links
// of one pose-set collide with each other?  FCL puts the L links of the pose-set in their own
// DynamicAABBTreeCollisionManager and self-collides it: every pair of links whose world AABBs overlap (closed
// intervals) goes through the same mesh-mesh narrowphase as an (environment object, link) pair.  Here the pair
// (a, b), a < b, is a candidate whose "environment object" is link a under its own pose: the kernel is launched
// with KernelParams::env = the agent's model pack and KernelParams::objects = one DevObject per link (model id
// only), so traverse_all and the exact pass of collide_device.cuh run unchanged.  The pose-set's bit is OR-ed
// into the verdict words the environment kernels wrote before (isInCollision returns true for either).
// Pair orientation: FCL's depends on the unspecified partition of std::nth_element inside its tree build; this
// path (and the oracle) use (lower link index, higher link index).
#include "collide_device.cuh"
#include "launch_cache.hpp"

namespace mptcc {
namespace {

__host__ __device__ constexpr unsigned self_warp_smem_bytes() { return warp_smem_bytes() + 6u * 32u * 8u; }

// Hit bookkeeping: one bit per pose-set slot of the warp's chunk; a candidate is (slot, link a) against the
// current link b.
struct SelfTrav {
	u64 hitmask, validmask;
	const double *chunk_poses;  // record of (slot s, link l): chunk_poses + (s * n_links + l) * rec_len
	unsigned rec_len, n_links, link_b;
	__device__ __forceinline__ bool dead(unsigned slot, const WarpMem &) const { return (hitmask >> slot) & 1ull; }
	__device__ __forceinline__ bool all_done() const { return hitmask == validmask; }
	__device__ __forceinline__ void mark_hit(unsigned slot, unsigned, const WarpMem &, const KernelParams &, int) {
		hitmask |= 1ull << slot;
	}
	__device__ __forceinline__ void mark_hits_per_lane(int slot, unsigned, const WarpMem &, int) {
		unsigned b = __ballot_sync(0xffffffffu, slot >= 0);
		while (b) {
			const int l = __ffs(b) - 1;
			b &= b - 1;
			hitmask |= 1ull << __shfl_sync(0xffffffffu, slot, l);
		}
	}
	__device__ __forceinline__ const double *record(unsigned slot, unsigned link) const {
		return chunk_poses + ((size_t)slot * n_links + link) * rec_len;
	}
	__device__ __forceinline__ unsigned object(unsigned cand, const WarpMem &wm) const { return wm.cand_ids[cand] & 0x3ffffffu; }
	// relativeTransform(pose of link a, pose of link b) in fp64 -> wm.rel; lanes 0..11 compute one element each
	__device__ __forceinline__ const double *exact_rel(const KernelParams &p, const WarpMem &wm, const DevObject &, unsigned slot,
	                                                   unsigned cand, int lane) const {
		__syncwarp();
		if (lane < 12) {
			DevObject oa;
			double R2[9], T2[3];
			Quat q;
			decode_pose(record(slot, object(cand, wm)), p.pose_format, oa.R, oa.T, q, false);
			decode_pose(record(slot, link_b), p.pose_format, R2, T2, q, false);
			wm.rel[lane] = rel_element(oa, R2, T2, lane);
		}
		__syncwarp();
		return wm.rel;
	}
};

template <bool COUNT>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) self_collide_kernel(const __grid_constant__ KernelParams p) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	unsigned char *base = smem_raw + (size_t)warp * self_warp_smem_bytes();
	WarpMem wm;
	wm.leafq = reinterpret_cast<uint2 *>(base);
	base += kLeafQueueCap * 8;
	wm.cand = reinterpret_cast<float4 *>(base);
	base += kCandCap * 64;
	wm.box = reinterpret_cast<double *>(base);  // world box of link b, component-major, one column per slot
	base += 6 * 32 * 8;
	wm.rel = reinterpret_cast<double *>(base);
	base += 16 * 8;
	wm.cand_ids = reinterpret_cast<uint32_t *>(base);
	base += kCandCap * 4;
	wm.clq = reinterpret_cast<uint32_t *>(base);
	base += kClusterQueueCap * 4;
	wm.donor = base;
	base += 32;
	double *box_a = reinterpret_cast<double *>(smem_raw + (size_t)warp * self_warp_smem_bytes() + warp_smem_bytes());

	const unsigned long long n_chunks = (p.n_pose_sets + 31ull) >> 5;
	const unsigned rec_len = pose_rec_len(p.pose_format);
	const unsigned lt = (1u << lane) - 1u;
	unsigned long long ctr[CTR_N];
#pragma unroll
	for (int i = 0; i < CTR_N; ++i) ctr[i] = 0;

	while (true) {
		unsigned chunk = 0;
		if (lane == 0) chunk = atomicAdd(p.work_counter, 1u);
		chunk = __shfl_sync(FULL, chunk, 0);
		if (chunk >= n_chunks) break;
		const unsigned long long pose0 = (unsigned long long)chunk << 5;
		const unsigned long long left = p.n_pose_sets - pose0;
		const unsigned nvalid = left < 32ull ? (unsigned)left : 32u;
		const bool valid = (unsigned)lane < nvalid;
		SelfTrav tv;
		tv.validmask = nvalid >= 32u ? 0xffffffffull : ((1ull << nvalid) - 1ull);
		// pose-sets the environment kernels already found in collision need no self test
		tv.hitmask = (u64)p.verdict_bits[pose0 >> 5] & tv.validmask;
		const u64 before = tv.hitmask;
		tv.chunk_poses = p.poses + pose0 * p.n_links * rec_len;
		tv.rec_len = rec_len;
		tv.n_links = p.n_links;

		for (unsigned b = 1; b < p.n_links && !tv.all_done(); ++b) {
			const DevModel &m2 = p.agent.models[p.link_model[b]];
			tv.link_b = b;
			float f3a[3], f3b[3];  // (the outward-rounded fp32 copies are not needed here)
			if (valid) pose_world_box(tv.record((unsigned)lane, b), p.pose_format, m2, wm.box + lane, f3a, f3b);
			int ncand = 0;
			for (unsigned a = 0; a < b && !tv.all_done(); ++a) {
				// ---- AABB::overlap(link a, link b) of the lane's own pose-set, exact fp64, closed intervals ----
				bool ov = false;
				if (valid && !tv.dead((unsigned)lane, wm)) {
					pose_world_box(tv.record((unsigned)lane, a), p.pose_format, p.agent.models[p.link_model[a]], box_a + lane, f3a,
					               f3b);
					ov = !((box_a[0 * 32 + lane] > wm.box[3 * 32 + lane]) | (box_a[1 * 32 + lane] > wm.box[4 * 32 + lane]) |
					       (box_a[2 * 32 + lane] > wm.box[5 * 32 + lane]) | (box_a[3 * 32 + lane] < wm.box[0 * 32 + lane]) |
					       (box_a[4 * 32 + lane] < wm.box[1 * 32 + lane]) | (box_a[5 * 32 + lane] < wm.box[2 * 32 + lane]));
				}
				const unsigned om = __ballot_sync(FULL, ov);
				if (om == 0u) continue;
				const int mine = ncand + __popc(om & lt);
				if (ov) {
					wm.cand_ids[mine] = ((unsigned)lane << 26) | a;
					// fp64 relativeTransform(pose a, pose b) -> fp32 record in the recentred frames
					DevObject oa;
					double Ra[9], Ta[3];
					Quat q;
					decode_pose(tv.record((unsigned)lane, a), p.pose_format, Ra, Ta, q, false);
#pragma unroll
					for (int k = 0; k < 9; ++k) oa.R[k] = Ra[k];
					oa.T[0] = Ta[0];
					oa.T[1] = Ta[1];
					oa.T[2] = Ta[2];
					oa.model = p.objects[a].model;
					float4 rec[4];
					make_candidate_record(p, tv.record((unsigned)lane, b), oa, m2, (unsigned)lane, rec);
#pragma unroll
					for (unsigned k = 0; k < 4; ++k) wm.cand[rec_at((unsigned)mine, k)] = rec[k];
				}
				ncand += __popc(om);
				if (COUNT) ctr[CTR_CAND] += (unsigned)__popc(om);
				__syncwarp();
				if (ncand > kCandCap - 32) {
					traverse_all<COUNT>(p, wm, m2, ncand, tv, lane, ctr);
					ncand = 0;
				}
			}
			if (ncand > 0 && !tv.all_done()) traverse_all<COUNT>(p, wm, m2, ncand, tv, lane, ctr);
			__syncwarp();
		}
		if (lane == 0 && tv.hitmask != before) atomicOr(p.verdict_bits + (pose0 >> 5), (uint32_t)tv.hitmask);
	}
	if (COUNT && lane == 0) {
#pragma unroll
		for (int i = 0; i < CTR_N; ++i)
			if (ctr[i]) atomicAdd(p.counters + i, ctr[i]);
	}
}

}  // namespace

// p.env must be the agent's pack and p.objects one DevObject per link (see the header of this file).
int launch_self_collide(const KernelParams &p, int grid_ctas, bool count, void *stream) {
	const size_t smem = (size_t)kWarpsPerCta * self_warp_smem_bytes();
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const void *k = count ? reinterpret_cast<const void *>(self_collide_kernel<true>)
	                      : reinterpret_cast<const void *>(self_collide_kernel<false>);
	const cudaError_t e = prepare_kernel(k, kWarpsPerCta * 32, smem, nullptr);
	if (e != cudaSuccess) return (int)e;
	if (count)
		self_collide_kernel<true><<<grid_ctas, kWarpsPerCta * 32, smem, s>>>(p);
	else
		self_collide_kernel<false><<<grid_ctas, kWarpsPerCta * 32, smem, s>>>(p);
	return (int)cudaGetLastError();
}

}  // namespace mptcc
