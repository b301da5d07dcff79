// device_layout.hpp -- HBM layout of the flattened scene, shared by the host flattener (scene.cpp)
// and the kernels (collide_kernel.cu).  See DESIGN.md "Data layout in HBM".
//
// The environment's models (one per object) and the agent's models (one per distinct link mesh) are
// two ModelPacks; inside a pack all models share the arrays and a model is a (node_base, tri_base)
// window.  Node and primitive ids on the traversal stack are model-local.
//
//   node32 : float4[4] per node, 64 B, fast-pass OBB in the model's RECENTRED frame (origin = the
//            model's local AABB centre):
//              [0] = (axis0.xyz, centre.x)  [1] = (axis1.xyz, centre.y)  [2] = (axis2.xyz, centre.z)
//              [3] = (extent.xyz rounded up and inflated by 1e-6, int-bits link)
//            link >= 0: first child (children are link, link+1); link < 0: leaf, primitive = -(link+1).
//            node_base is odd, so every sibling pair (2k+1, 2k+2) shares one 128-byte line.
//   nodeq  : uint4[2] per node, 32 B, same numbering as node32: the optional traversal node of the np2/np3 kernels
//            (qnode = 1) -- ONE 256-bit load per node, a sibling pair is 64 bytes.
//              [0].x, .y, .z = centre (fp32)      [0].w = extent.x | extent.y << 16      [1].x = extent.z | link[15:0] << 16
//                      (extents binary16, rounded UP after the widening below)
//              [1].y, .z, .w = 96 bits: the orientation as a unit quaternion (w, x, y, z), four 22-bit fixed-point
//                      components Q (value Q * 2^-21 - 1) in bits [0,22) [22,44) [44,66) [66,88), then link[23:16]
//            The rows of the quaternion's rotation matrix are the box axes (a left-handed triple is stored with axis 2
//            negated: the same box).  link: the node32 link as a 24-bit two's-complement number (one PRMT sign-extends it).  A decoded
//            axis component is within 4e-6 of the true one; the extents are widened by sqrt(3) * 4e-6 * (sum of
//            extents) and (1 + 1e-6) so that the decoded box contains FCL's box.
//   tri32  : float4[3] per triangle (recentred corner xyz, w unused), indexed by primitive id.
//   node64 : double[16] per node: FCL's OBB verbatim -- axis (9), To (3), extent (3), size =
//            extent.sqrLength() (1).  Only read when a positive leaf pair replays FCL's path.
//   anc    : uint32[depth] per triangle: the root-to-leaf chain of node ids that ends at the leaf holding
//            that primitive (unused levels = 0xffffffff); depth = DevModel::depth.  Lets the exact pass
//            fetch a whole ancestor chain with one coalesced load instead of walking the tree.
//   tri64  : double[9] per triangle, model coordinates verbatim (exact pass).
//
// Broadphase tree (bp_nodes): the objects are sorted along a Morton curve and cut into clusters of
// kClusterSize consecutive objects; a binary tree over the cluster sequence is stored in pre-order
// ("threaded": no stack needed).  Node n = one uint4 (16 B, one load per visit):
//   x = qmin.x | qmin.y << 16,  y = qmin.z | qmax.x << 16,  z = qmax.y | qmax.z << 16
//       the node box on a 16-bit grid over the scene box, q(x) = (x - bp_origin) * bp_inv_cell, minima rounded
//       down and maxima up (the scene spans 65 000 cells);
//   w = inner node: escape = index of the next node once this subtree is done (n_bp_nodes = end), the first
//       child is node n+1;  leaf: 0x80000000 | cluster, for objects [cluster*kClusterSize, +kClusterSize)
//       (a leaf's escape is n+1).
// A pose box is moved to the same grid with one more cell of slack (quantize_pose_box).
// The grid boxes are conservative; every object of a touched cluster is re-tested in fp64.
#pragma once

#include <cstdint>

namespace mptcc {

// Doubles per pose record of each format (include/mptcc.h): RT12, TQ7, T3, XYZCS.
#ifdef __CUDACC__
__host__ __device__
#endif
inline constexpr unsigned pose_rec_len(unsigned fmt) { return fmt == 1u ? 7u : fmt == 2u ? 3u : fmt == 3u ? 5u : 12u; }

// Broadphase grid (single-link agents, split pipeline).  The scene box is cut into grid_n[0] x [1] x [2] cells no
// smaller than the largest world box a pose of the agent can have, so a pose box touches at most two cells per axis.
// Cell (i, j, k) owns the list of the objects whose box touches one of the cells [i, i+1] x [j, j+1] x [k, k+1]:
// grid_cells[(i n1 + j) n2 + k] = (first entry, count), entries = 64 bytes: the object's exact fp64 world box
// (min xyz, max xyz, what obj_aabb holds), the object index (as the low word of the 7th double) and padding.
// grid_cell() is the one cell-index function of the host (objects) and the device (poses); it is monotone in x, so
// two boxes that overlap always share a cell along every axis.
struct GridEntry {
	double aabb[6];
	unsigned long long obj;
	unsigned long long pad;
};
#ifdef __CUDACC__
__host__ __device__
#endif
inline int grid_cell(double x, float origin, float inv, int n) {
	const float v = ((float)x - origin) * inv;
	return v >= (float)n ? n - 1 : (v > 0.0f ? (int)v : 0);
}

struct DevModel {
	uint32_t node_base;
	uint32_t tri_base;
	uint32_t num_tris;
	uint32_t depth;
	double ctr[3];         // recentring origin (fp64)
	double aabb_min[3];    // BVHModel::computeLocalAABB (model frame, not recentred)
	double aabb_max[3];
	double aabb_center[3];
	double aabb_radius;
	float r2;              // max |v - ctr|_2 over the vertices, rounded up
	float rinf;            // max |v - ctr|_inf over the vertices, rounded up
	unsigned long long anc_base;  // first word of this model's ancestor chains in ModelPack::anc
	// replay certificate (collide_kernel.cu, "path certificate"):
	double cmax1;  // max over nodes of |To|_1
	double esum;   // max over nodes of extent[0] + extent[1] + extent[2]
};

struct DevObject {
	double R[9];     // tf1 rotation (row-major)
	double T[3];     // tf1 translation
	double aabb[6];  // world AABB: min xyz, max xyz (CollisionObject::computeAABB)
	uint32_t model;
	uint32_t identity;  // 1: R is exactly the identity and T exactly zero (environment vertices given in world coordinates)
};

// One set of flattened models (the environment's, or the agent's).
struct ModelPack {
	const float4 *node32;
	const float4 *tri32;
	const double *node64;
	const double *tri64;
	const uint32_t *anc;
	const float *mu_chain;  // per triangle: min over its ancestor chain (leaf included) of the sum of the two
	                        // smallest OBB extents of the node, rounded down (path certificate)
	const DevModel *models;
	const uint4 *nodeq;     // 2 x uint4 per node: the quantised fast-pass node of the np2 kernel (see above)
};

struct KernelParams {
	ModelPack env;    // models referenced by DevObject::model
	ModelPack agent;  // models referenced by link_model[]
	const DevObject *objects;    // in broadphase (Morton-cluster) order
	const double *obj_aabb;      // [6][n_objects_pad] world AABB min xyz, max xyz (exact fp64), component-major,
	                             // same order, padded with empty boxes to a multiple of kClusterSize
	uint32_t n_objects_pad;
	const uint4 *bp_nodes;       // broadphase tree, one uint4 per node (see above)
	uint32_t n_bp_nodes;
	float bp_origin[3], bp_inv_cell[3];
	const uint32_t *link_model;  // [n_links] agent-pack model id of link j
	uint32_t n_objects;
	uint32_t n_links;
	const double *poses;
	uint32_t pose_format;
	uint32_t ppw_log2;  // poses per warp = 1 << ppw_log2
	unsigned long long n_pose_sets;
	uint32_t *verdict_bits;
	uint32_t *work_counter;
	// split pipeline (split_kernels.cu): candidate list written by the broadphase kernel
	float4 *cand_rec;               // [cand_cap][4] fp32 candidate records
	double *cand_rel;               // [cand_cap][12] fp64 relativeTransform (R, T) of the candidate (exact pass)
	uint32_t *cand_obj;             // [cand_cap] environment object of the candidate
	uint32_t *cand_pose;            // [cand_cap] pose-set index of the candidate
	unsigned int *cand_total;       // [1] candidates appended (can exceed cand_cap: then *overflow = 1)
	unsigned int *overflow;         // [1]
	unsigned int cand_cap;
	const unsigned int *run_if;     // fused kernel: return at once unless *run_if != 0 (nullptr: always run)
	uint32_t *work_counter2;        // batch counter of the narrowphase kernel
	unsigned long long *counters;  // [CTR_N] or nullptr
	uint32_t near_first;           // 1: of two overlapping children, descend the closer one first
	uint32_t steal_bottom;         // 1: thieves take the bottom (largest, most speculative) pending pair
	uint32_t no_steal;             // 1: lanes never steal pending subtrees (experiment knob)
	uint32_t always_replay;        // 1: replay FCL's path for every positive pair (ignore the certificate)
	// second-generation narrowphase (np2_kernels.cu)
	uint2 *cand_base;              // [cand_cap] node_base, tri_base of the candidate's environment model
	uint4 *xq;                     // [xq_cap] exact-pass queue: candidate list index, env prim, agent prim, certain-contact flag
	unsigned int *xq_total;        // [1] entries appended (can exceed xq_cap: then *overflow = 1)
	unsigned int xq_cap;
	uint32_t expand4;              // 1: a narrowphase step expands both trees (four tests), 0: FCL's one-tree descent (two)
	uint32_t qnode;                // 1: traverse nodeq (32-byte nodes), 0: node32
	uint32_t agent_top;            // agent nodes [0, agent_top) of link model 0 are staged in shared memory (breadth-first ids)
	// third-generation narrowphase (np3_kernels.cu)
	const uint2 *grid_cells;       // broadphase grid (see above); nullptr: walk the cluster tree
	const GridEntry *grid_entries;
	float grid_origin[3], grid_inv[3];
	int grid_n[3];
	float *cand_cert;              // [cand_cap] path-certificate threshold of the candidate: a positive leaf pair needs no path
	                               // replay if the mu_chain of both triangles is >= this
	const float4 *agent_swz;       // link model 0's node32 records in the bank-swizzled order the np3 kernel stages with TMA
};

// index of the counters in KernelParams::counters
enum { CTR_CAND = 0, CTR_BV = 1, CTR_LEAF = 2, CTR_EXACT = 3, CTR_PATH = 4, CTR_N = 5 };

#ifndef MPTCC_CLUSTER
#define MPTCC_CLUSTER 8
#endif
constexpr int kClusterSize = MPTCC_CLUSTER;
constexpr int kWarpsPerCta = 8;
constexpr int kLeafQueueCap = 96;   // a node step appends <= 64 pairs; the leaf phase runs at >= 32
constexpr int kClusterQueueCap = 64;  // (pose, cluster) entries waiting for the exact AABB test
constexpr unsigned kAgentTopMaxQ = 1023; // the same with 32-byte quantised nodes: 10 full levels, 32 KB
constexpr unsigned kAgentTopMax = 511;  // top of the (breadth-first numbered) agent tree staged in shared memory: 9 full levels, 32 KB
constexpr int kLaneStackCap = 64;     // per-lane depth-first stack (local memory); needs depth1 + depth2 + 2 <= 64
#ifndef MPTCC_CAND_CAP
#define MPTCC_CAND_CAP 96
#endif
constexpr int kCandCap = MPTCC_CAND_CAP;          // candidate table: (pose, object) pairs traversed together (<= 256)
// per-warp shared memory: leaf queue + candidate table + pose boxes + fp64 relative transform +
// candidate ids + cluster queue
#ifdef __CUDACC__
__host__ __device__
#endif
inline constexpr unsigned warp_smem_bytes() {
	return kLeafQueueCap * 8u + kCandCap * 64u + 6u * 32u * 8u + 16u * 8u + kCandCap * 4u +
	       kClusterQueueCap * 4u + 32u;
}

// Launch wrapper implemented in collide_kernel.cu.  Returns a cudaError_t as int.
int launch_collide(const KernelParams &p, int grid_ctas, bool count, void *stream);
int collide_max_ctas_per_sm(bool count);

// split_kernels.cu: broadphase kernel (poses -> candidate list) and narrowphase kernel (candidate list ->
// verdict bits); single-link agents.
int launch_broadphase(const KernelParams &p, int sm_count, bool count, bool np2, void *stream);
// self_kernel.cu: link-vs-link test of every pose-set (checkSelfCollision); p.env = the agent's pack,
// p.objects = one DevObject per link.  ORs into verdict_bits.
int launch_self_collide(const KernelParams &p, int grid_ctas, bool count, void *stream);
int launch_narrowphase(const KernelParams &p, int sm_count, bool count, bool long_batches, void *stream);
// np2_kernels.cu: second-generation narrowphase (TMA-staged agent tree top, child-frame SAT) + its exact-pass kernel
int launch_narrowphase2(const KernelParams &p, int sm_count, bool count, bool long_batches, void *stream);
int launch_exact(const KernelParams &p, int sm_count, bool count, void *stream);
// np3_kernels.cu: third-generation narrowphase (streaming lanes, whole agent tree in shared memory); uses the np2
// broadphase records and exact-pass kernel
int launch_narrowphase3(const KernelParams &p, int sm_count, bool count, void *stream);
unsigned np3_agent_top_max();

// edge_kernels.cu
int launch_expand_omni_edges(const double *d_edges, const unsigned long long *d_offsets, unsigned long long n_edges,
                             unsigned long long total_poses, double dt, double *d_poses, void *stream);
int launch_reduce_segments(const uint32_t *d_bits, const unsigned long long *d_offsets, unsigned long long n_segments,
                           unsigned char *d_safe, void *stream);

}  // namespace mptcc
