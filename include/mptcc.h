/*
 * mptcc.h -- C ABI of libmptcc.so: the B200 (sm_100a) batched mesh-vs-mesh collision checker that
 * replaces the FCL-backed validity test of csbence/motionplanningtoolkit.
 *
 * Drop-in boundary.  The reference has no FFI; its innermost seam for this path is the static C++
 * function
 *     MeshHandler::isInCollision(const StaticEnvironmentMeshHandler&,
 *                                const std::vector<const SimpleAgentMeshHandler*>&,
 *                                const std::vector<std::vector<fcl::Transform3f>>& poses,
 *                                bool checkSelfCollision = false, bool debug = false)
 * (/root/reference/utilities/meshhandler.hpp:187-191), reached from Map3D::safeEdge / safePoses
 * (/root/reference/workspaces/map3d.hpp:33-42).  The entry points below are what a binding of that
 * seam needs; the headers under include/mpt_b200/ are the C++ shim that keeps the reference's class and method
 * names on top of them.
 *
 * Conventions: plain C, opaque context, integer status (0 = OK), no exceptions, no C++/torch types.
 * The caller owns every buffer it passes; inputs of set_* are copied.  One context is not re-entrant
 * (calls on one ctx are serialised by the caller); distinct contexts are independent.  There is no
 * CPU fallback: every check_* call fails with MPTCC_ERR_CUDA / MPTCC_ERR_NO_DEVICE when no sm_100
 * device is usable.
 */
#ifndef MPTCC_H
#define MPTCC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPTCC_VERSION 1

/* status codes */
#define MPTCC_OK 0
#define MPTCC_ERR_INVALID 1   /* bad argument                                   */
#define MPTCC_ERR_STATE 2     /* environment / agent not set                    */
#define MPTCC_ERR_NO_DEVICE 3 /* no CUDA device / not an sm_100 device          */
#define MPTCC_ERR_CUDA 4      /* a CUDA runtime call failed (see last_error)    */
#define MPTCC_ERR_NOMEM 5
#define MPTCC_ERR_LIMIT 6     /* tree too deep / too many nodes for the kernel  */

/* Pose record formats.  A pose-set is n_links consecutive records.
 *  RT12: 12 doubles, R row-major (9) then T (3): fcl::Transform3f(R,T) and fcl::Transform3f(T)
 *        (identity R) -- what the agents' getPoses produce, e.g.
 *        /root/reference/agents/omnidirectional.hpp:224-242, agents/snake_trailers.hpp:432,455.
 *  TQ7 : 7 doubles, x y z qw qx qy qz: fcl::Transform3f(fcl::Quaternion3f(w,x,y,z), fcl::Vec3f)
 *        -- fcl_helpers::parseTransform (/root/reference/utilities/fcl_helpers.hpp:16-25) and the
 *        PRMLite samples (/root/reference/discretizations/workspace/prmlite.hpp:246-258).
 *  T3  : 3 doubles, x y z: fcl::Transform3f(fcl::Vec3f) -- identity rotation AND identity quaternion, i.e. the
 *        translated tight AABB in the broadphase; what Omnidirectional::getPoses emits
 *        (/root/reference/agents/omnidirectional.hpp:224-242).  Same verdicts as RT12 with R = I, a quarter of the bytes.
 *  XYZCS: 5 doubles, x y z c s: fcl::Transform3f(R, T) with the planar rotation the blimp and the snake build,
 *        R(0,0)=c R(1,0)=-s R(0,1)=s R(1,1)=c R(2,2)=1 (/root/reference/agents/blimp.hpp:345-355,
 *        agents/snake_trailers.hpp:422-432,445-455); c = cos(theta), s = sin(theta) are evaluated by the CALLER (host
 *        libm, as the reference does), so the verdicts are those of the RT12 record with the same entries. */
#define MPTCC_POSE_RT12 0
#define MPTCC_POSE_TQ7 1
#define MPTCC_POSE_T3 2
#define MPTCC_POSE_XYZCS 3

typedef struct mptcc_ctx mptcc_ctx;

/* One sub-model as handed to fcl::BVHModel::addSubModel(verts, tris)
 * (/root/reference/utilities/meshhandler.hpp:44-47,130-133). */
typedef struct mptcc_mesh {
	const double *xyz;   /* [nv][3] */
	uint32_t nv;
	const uint32_t *tri; /* [nt][3] vertex indices */
	uint32_t nt;
} mptcc_mesh;

typedef struct mptcc_counters {
	uint64_t n_pose_sets;   /* pose-sets checked since the counters were last reset                */
	uint64_t n_candidates;  /* (pose, link, env object) pairs whose world AABBs overlap            */
	uint64_t n_bv_tests;    /* fp32 OBB-OBB tests executed by the traversal                        */
	uint64_t n_leaf_tests;  /* fp32 triangle-triangle filter tests                                 */
	uint64_t n_exact_tri;   /* leaf pairs re-decided in fp64 (inside the fp32 error band)          */
	uint64_t n_path_checks; /* fp64 replays of the reference traversal path for a positive pair    */
	uint64_t n_kernel_launches;
} mptcc_counters;

/* ---- context ------------------------------------------------------------------------------ */
/* n_devices GPUs (1/2/4/8); device_ids may be NULL for 0..n_devices-1. */
int mptcc_create(mptcc_ctx **out, int n_devices, const int *device_ids);
void mptcc_destroy(mptcc_ctx *ctx);
const char *mptcc_last_error(const mptcc_ctx *ctx);
int mptcc_num_devices(const mptcc_ctx *ctx);

/* ---- scene -------------------------------------------------------------------------------- */
/* StaticEnvironmentMeshHandler(file, pose) (/root/reference/utilities/meshhandler.hpp:18-55): one
 * collision object per non-empty sub-mesh.  obj_pose7: [n_objects][7] x y z qw qx qy qz (the
 * reference gives every object the same "Environment Location").  Builds the FCL-identical OBB
 * trees on the host, flattens them and replicates them on every device of the context.
 * Both set_* calls first wait for everything outstanding on the context (submits, device-buffer launches): the
 * scene arrays are replaced in place.  A mesh with a triangle index >= nv is refused (MPTCC_ERR_INVALID). */
int mptcc_set_environment(mptcc_ctx *ctx, int n_objects, const mptcc_mesh *meshes, const double *obj_pose7);

/* Agent::getMeshes() (/root/reference/agents/omnidirectional.hpp:186-189,
 * agents/snake_trailers.hpp:215-218): n_links models, link j of every pose-set uses link_meshes[j].
 * Meshes with identical (xyz, tri) pointers share one device model. */
int mptcc_set_agent(mptcc_ctx *ctx, int n_links, const mptcc_mesh *link_meshes);

/* ---- queries: host buffers ---------------------------------------------------------------- */
/* One MeshHandler::isInCollision per pose-set.  poses: [n_pose_sets][n_links] records of
 * pose_format; verdict_bits: ceil(n/32) words, bit i%32 of word i/32 = 1 <=> pose-set i collides.
 * Pose-sets are split in contiguous slices over the context's devices (sub-batches ping-pong over two streams so
 * that copies overlap kernels); a call with at most 256 KB of poses -- one edge -- is served without any copy
 * from pinned, device-mapped staging memory on the first device. */
int mptcc_check_poses(mptcc_ctx *ctx, uint64_t n_pose_sets, const double *poses, int pose_format,
                      uint32_t *verdict_bits);

/* The same call split in two, for callers that stream batches (a planner that prepares slice k+1 while slice
 * k is being checked): submit enqueues the copies and kernels and returns a ticket at once; wait blocks until
 * that ticket (and every earlier one) is complete, ticket 0 = everything submitted.  poses and verdict_bits
 * must stay valid until the wait returns and should be pinned (mptcc_alloc_pinned), otherwise the copies
 * are staged synchronously by the driver and submit does not overlap anything.  Up to 4 submits may be
 * outstanding; a further submit first waits for the oldest.  mptcc_check_poses == submit + wait. */
int mptcc_submit_poses(mptcc_ctx *ctx, uint64_t n_pose_sets, const double *poses, int pose_format,
                       uint32_t *verdict_bits, uint64_t *ticket);
int mptcc_wait(mptcc_ctx *ctx, uint64_t ticket);

/* Map3D::safeEdge semantics for many edges at once (/root/reference/workspaces/map3d.hpp:33-37):
 * segment s owns pose-sets [seg_begin[s], seg_begin[s+1]); segment_safe[s] = 1 iff none collides
 * (an empty segment is safe). */
int mptcc_check_segments(mptcc_ctx *ctx, uint64_t n_segments, const uint64_t *seg_begin, const double *poses,
                         int pose_format, uint8_t *segment_safe);

/* Omnidirectional::getPoses on the device (/root/reference/agents/omnidirectional.hpp:202-247):
 * edge e = start xyz, end xyz (6 doubles); its pose-sets are the reference's translation-only interpolation
 * at spacing dt (iterations = (unsigned)(dist/dt); {start,end} if 0; the end pose is appended iff
 * iterations*dt < dist), generated on the GPU with the reference's fp64 arithmetic.
 * edge_safe[e] = 1 iff none collides (Map3D::safeEdge).  Needs a single-link agent. */
int mptcc_check_omni_edges(mptcc_ctx *ctx, uint64_t n_edges, const double *start_end, double dt, uint8_t *edge_safe);

/* ---- queries: device buffers (inputs already resident in HBM) -------------------------------- */
/* d_poses / d_verdict_bits are device pointers on device slot `slot` (0..n_devices-1); the launch is
 * enqueued on `cuda_stream` (a cudaStream_t, NULL = the context's stream) and NOT synchronised.
 * Every launch of a slot shares that slot's candidate list and work counters: keep at most ONE launch in flight per
 * slot (chain launches on one stream, or synchronise in between); launches on different slots are independent. */
int mptcc_check_poses_device(mptcc_ctx *ctx, int slot, uint64_t n_pose_sets, const double *d_poses,
                             int pose_format, uint32_t *d_verdict_bits, void *cuda_stream);
int mptcc_synchronize(mptcc_ctx *ctx);

/* Pinned host memory for pose / verdict buffers (full PCIe rate for the host-buffer calls). */
void *mptcc_alloc_pinned(size_t bytes);
void mptcc_free_pinned(void *p);

/* ---- instrumentation ------------------------------------------------------------------------- */
/* enable != 0: later launches use the counting kernel variant (slower). */
int mptcc_enable_counters(mptcc_ctx *ctx, int enable);
int mptcc_get_counters(mptcc_ctx *ctx, mptcc_counters *out);
int mptcc_reset_counters(mptcc_ctx *ctx);
/* Device milliseconds of the last host-buffer check call on device slot `slot`: from the first kernel to the end
 * of the last kernel (CUDA events on the launching streams; copies that overlap the kernels are inside). */
int mptcc_last_kernel_ms(mptcc_ctx *ctx, int slot, float *ms);
/* With option "time_phases" = 1: device milliseconds of the four kernels of the last split-pipeline launch
 * on `slot` (CUDA events on the launching stream): ms4 = {broadphase, narrowphase, exact pass, overflow fall-back}
 * (the round-1 narrowphase, options "np3" = 0 and "np2" = 0, has no separate exact pass: ms4[2] is then ~0). */
int mptcc_last_phase_ms(mptcc_ctx *ctx, int slot, float *ms4);
/* Behaviour: "self_collision" (0/1, default 0) = the checkSelfCollision argument of
 * MeshHandler::isInCollision (/root/reference/utilities/meshhandler.hpp:187-219): every later check also
 * tests the links of each pose-set against each other (pairs with overlapping world AABBs, same mesh-mesh
 * narrowphase) and reports self-collision OR environment collision.
 * Tuning knobs: "poses_per_warp" (0 = auto, power of two <= 64), "ctas_per_sm", "sub_batch", "split" (0/1),
 * "split_min", "split_cap", "long_batches" (-1 = by launch size), "always_replay", "time_phases", "near_first",
 * "no_steal", "steal_bottom"; kernel selection of the split pipeline (large single-link batches): "grid" (1 = grid
 * broadphase, the default; 0 = cluster-tree walk), "np3" (1 = streaming narrowphase, the default; 0 = the round-1
 * kernel), "np2" (1 = second-generation narrowphase; selects it instead of np3), "qnode" (1 = 32-byte quantised
 * traversal nodes, the default; 0 = 64-byte fp32 nodes), "agent_top" (fp32 nodes only: agent tree nodes staged in
 * shared memory by TMA; np3: 0 / -1 = none, np2: -1 = as many as fit), "l2_persist" (1 = the narrowphase runs with a
 * persisting-L2 access-policy window over the environment's traversal nodes; the window is set on the launch stream
 * before the narrowphase and cleared after it, so a caller-supplied stream loses a window of its own; 0 = off; -1, the
 * default = on for agents of >= 128 triangles against >= 16 MB of environment nodes, where the narrowphase dominates).  Every combination gives the same verdicts.
 * The environment variable MPTCC_OPTIONS="name=value,..." applies options to every new context. */
int mptcc_set_option(mptcc_ctx *ctx, const char *name, long value);
/* What the current scene and options resolve to (tests and tools): "qnode_active" (1: the split pipeline traverses the
 * 32-byte quantised nodes; 0: a pack could not be quantised -- an extent above 65 000, a model beyond 2^22 triangles --
 * or the option is off), "grid_cells" / "grid_entries" (size of the broadphase grid; 0: the cluster tree is walked),
 * "agent_nodes", "env_nodes". */
int mptcc_get_info(mptcc_ctx *ctx, const char *name, long *value);

/* ---- host-side scene introspection (no GPU needed; used by the CPU test-suite) --------------- */
/* A context WITHOUT devices: mptcc_set_environment / mptcc_set_agent build the host side of the scene as usual (OBB
 * trees, object boxes, the broadphase cluster tree and grid); every check call fails with MPTCC_ERR_NO_DEVICE. */
int mptcc_debug_create_host_only(mptcc_ctx **out);
/* What the broadphase structures answer for one pose box (box6 = min xyz, max xyz), looked up the way the kernels do:
 * *n_listed = objects the structure lists for the box (the grid cell's entries / the objects of the clusters the tree
 * walk reaches); return value = how many of those overlap the box under fcl::AABB::overlap (closed intervals), their
 * world boxes written to aabbs (6 doubles each, at most cap).  -1: the structure does not exist for this scene
 * (multi-link agent or non-finite boxes: no grid); -2 (grid): the box cannot be looked up (not finite, or wider than
 * a cell) -- the kernels then fall back to the tree. */
int mptcc_debug_grid_lookup(mptcc_ctx *ctx, const double *box6, double *aabbs, uint32_t cap, uint32_t *n_listed);
int mptcc_debug_tree_lookup(mptcc_ctx *ctx, const double *box6, double *aabbs, uint32_t cap, uint32_t *n_listed);

/* ---- host-side BVH introspection (no GPU needed; used by the CPU test-suite) ----------------- */
typedef struct mptcc_bvh mptcc_bvh;
int mptcc_bvh_build(const mptcc_mesh *mesh, mptcc_bvh **out);
void mptcc_bvh_free(mptcc_bvh *b);
uint32_t mptcc_bvh_num_nodes(const mptcc_bvh *b);
uint32_t mptcc_bvh_depth(const mptcc_bvh *b);
/* axis [n][9], To [n][3], extent [n][3], first_child [n], first_primitive [n], num_primitives [n];
 * any pointer may be NULL. */
void mptcc_bvh_get_nodes(const mptcc_bvh *b, double *axis, double *To, double *extent, int32_t *first_child,
                         int32_t *first_primitive, int32_t *num_primitives);
void mptcc_bvh_get_primitive_indices(const mptcc_bvh *b, uint32_t *out);
/* The 32-byte quantised traversal nodes the device would get for this model (csrc/device_layout.hpp, nodeq):
 * words [n][8]; device_id [n] (may be NULL) = device number of FCL node i (node n of `words` is device node n);
 * ctr3 (may be NULL) = the model's recentring origin the stored centres are relative to.  Returns 1 if every node
 * could be quantised within the bounds of DESIGN.md section 6, 0 if the pack would fall back to fp32 nodes, -1 on
 * bad arguments. */
int mptcc_bvh_get_quantised_nodes(const mptcc_bvh *b, int breadth_first, uint32_t *words, uint32_t *device_id, double *ctr3);
/* The arrays of the exact pass for this model (csrc/device_layout.hpp): node64 [n][16] (FCL's OBB verbatim + size, device
 * numbering) and anc [num_tris][*depth] (root-to-leaf node chain per triangle, unused levels 0xffffffff); either may be
 * NULL.  Returns 0, -1 on bad arguments. */
int mptcc_bvh_get_exact_arrays(const mptcc_bvh *b, int breadth_first, double *node64, uint32_t *anc, uint32_t *depth);
/* What the path certificate uses of this model: mu_chain [num_tris] (may be NULL) and model6 = the
 * recentring origin (3), cmax1 = max |To|_1 over the nodes, esum = max sum of extents over the nodes, r2. */
int mptcc_bvh_get_certificate_data(const mptcc_bvh *b, float *mu_chain, double *model6);
/* aabb6 = min xyz, max xyz */
void mptcc_bvh_get_local_aabb(const mptcc_bvh *b, double *aabb6, double *center3, double *radius);

#ifdef __cplusplus
}
#endif
#endif /* MPTCC_H */
