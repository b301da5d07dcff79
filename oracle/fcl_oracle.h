/*
 * fcl_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C, fp64 restatement of the collision path the reference toolkit runs through
 * FCL 0.3.2: MeshHandler::isInCollision (/root/reference/utilities/meshhandler.hpp:187-243)
 * over fcl::BVHModel<fcl::OBBRSS>, fcl::CollisionObject, fcl::DynamicAABBTreeCollisionManager
 * and fcl::collide (/root/reference/utilities/fcl_helpers.hpp:52-65).
 *
 * PARITY UNPINNED: FCL itself is a third-party dependency that is not vendored in the
 * reference ("version 0.3.2", /root/reference/README.md:13; pkg_search_module(FCL REQUIRED fcl),
 * /root/reference/CMakeLists.txt:20) and is not installed in this image; the reference has no
 * tests or golden vectors.  The arithmetic below restates FCL 0.3.x's published algorithm
 * ([FCL-recalled]) operation by operation; it is pinned only by analytic
 * known-answer tests and by a structure-free brute-force cross-check (tests/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this file.  Build: gcc -O2 -ffp-contract=off (no -march=native, no fast-math).
 */
#ifndef FCL_ORACLE_H
#define FCL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Pose record formats (mirrors include/mptcc.h). */
#define ORC_POSE_RT12 0 /* 12 doubles: R row-major (9), T (3)  == fcl::Transform3f(R,T) / Transform3f(T) */
#define ORC_POSE_TQ7 1  /* 7 doubles: x y z qw qx qy qz        == fcl::Transform3f(Quaternion3f(w,x,y,z), T) */

typedef struct orc_model orc_model; /* fcl::BVHModel<fcl::OBBRSS> after beginModel/addSubModel/endModel */
typedef struct orc_env orc_env;     /* StaticEnvironmentMeshHandler: E objects + broadphase tree      */

typedef struct orc_counters {
	uint64_t n_pose_sets;  /* isInCollision-equivalents evaluated                                  */
	uint64_t n_aabb_tests; /* broadphase AABB-AABB overlap tests (dual tree recursion)             */
	uint64_t n_obj_pairs;  /* defaultCollisionFunction callbacks that ran fcl::collide             */
	uint64_t n_bv_tests;   /* OBB-OBB tests (BVTesting)                                            */
	uint64_t n_leaf_tests; /* triangle-triangle tests (leafTesting)                                */
} orc_counters;

/* ---- model (BVH) -------------------------------------------------------------------------- */
orc_model *orc_model_build(const double *xyz, uint32_t nv, const uint32_t *tri, uint32_t nt);
void orc_model_free(orc_model *m);
uint32_t orc_model_num_nodes(const orc_model *m);
uint32_t orc_model_num_tris(const orc_model *m);
/* Node dump: axis[n][9] (axis[k] = 3 consecutive doubles), To[n][3], extent[n][3],
 * first_child[n], first_primitive[n], num_primitives[n]; any pointer may be NULL. */
void orc_model_get_nodes(const orc_model *m, double *axis, double *To, double *extent, int32_t *first_child,
                         int32_t *first_primitive, int32_t *num_primitives);
void orc_model_get_primitive_indices(const orc_model *m, uint32_t *out);
/* aabb6 = min xyz, max xyz; center3; radius */
void orc_model_get_local_aabb(const orc_model *m, double *aabb6, double *center3, double *radius);

/* ---- environment -------------------------------------------------------------------------- */
/* pose7: [n_objects][7] = x y z qw qx qy qz (the "Environment Location" layout parsed by
 * fcl_helpers::parseTransform, /root/reference/utilities/fcl_helpers.hpp:16-25). The models are
 * borrowed (must outlive the env). */
orc_env *orc_env_create(int n_objects, orc_model *const *models, const double *pose7);
void orc_env_free(orc_env *e);
void orc_env_get_object_aabb(const orc_env *e, int obj, double *aabb6);

/* ---- queries ------------------------------------------------------------------------------- */
/* MeshHandler::isInCollision(env, agent, poses[P][L]) : one bool for the whole batch, stops at the
 * first contact.  poses: [P][L] records of `pose_format`.  counters may be NULL (accumulated). */
int orc_is_in_collision(const orc_env *env, int n_links, orc_model *const *links, uint64_t n_pose_sets,
                        const double *poses, int pose_format, orc_counters *counters);

/* One isInCollision per pose-set -> bit i of verdict_bits (bit = 1: in collision).
 * n_threads > 1 splits the pose range over that many pthreads (word-aligned slices). */
int orc_verdicts(const orc_env *env, int n_links, orc_model *const *links, uint64_t n_pose_sets,
                 const double *poses, int pose_format, uint32_t *verdict_bits, int n_threads,
                 orc_counters *counters);

/* The same with checkSelfCollision = true (meshhandler.hpp:205-219): bit = the links of the pose-set collide
 * with each other OR with the environment.  orc_self_collision is the per-pose-set self test alone. */
int orc_verdicts_self(const orc_env *env, int n_links, orc_model *const *links, uint64_t n_pose_sets,
                      const double *poses, int pose_format, uint32_t *verdict_bits, int n_threads, orc_counters *counters);
int orc_self_collision(int n_links, orc_model *const *links, const double *pose_set, int pose_format,
                       orc_counters *counters);

/* Structure-free second oracle: verdict = exists object, link with overlapping world AABBs and a
 * triangle pair for which Intersect::intersect_Triangle(...,R,T) holds.  No BVH, no OBB tests. */
int orc_verdicts_bruteforce(const orc_env *env, int n_links, orc_model *const *links, uint64_t n_pose_sets,
                            const double *poses, int pose_format, uint32_t *verdict_bits);

/* The same verdicts, fast enough for deep trees and large scenes: pose range over n_threads pthreads, and a pair of
 * triangles is only tested when their axis-aligned boxes (grown by 1e-9 of the coordinate magnitude) overlap.  Still
 * no BVH, no OBB fit, no OBB test: independent of everything the BVH oracle could get wrong there. */
int orc_verdicts_bruteforce_fast(const orc_env *env, int n_links, orc_model *const *links, uint64_t n_pose_sets,
                                 const double *poses, int pose_format, uint32_t *verdict_bits, int n_threads);

/* ---- primitives exposed for unit tests ------------------------------------------------------ */
/* OBB overlap(R0,T0,b1,b2) -> 1 if NOT disjoint. b = axis[9], To[3], extent[3] */
int orc_obb_overlap(const double *R0, const double *T0, const double *axis1, const double *To1,
                    const double *ext1, const double *axis2, const double *To2, const double *ext2);
/* Intersect::intersect_Triangle(P1,P2,P3,Q1,Q2,Q3,R,T) */
int orc_tri_intersect(const double *P /*[9]*/, const double *Q /*[9]*/, const double *R, const double *T);
/* Transform helpers: fill R[9], T[3], q[4] from a pose record. */
void orc_pose_decode(const double *rec, int pose_format, double *R, double *T, double *q);
/* 1 iff collisionRecurse on (m1, m2) under x1 = R x2 + T reaches the leaf pair of triangles (prim1, prim2) */
int orc_path_reaches(const orc_model *m1, const orc_model *m2, const double *R, const double *T, int prim1, int prim2);
/* fcl::relativeTransform(tf1, tf2) of two pose records: R = R1^T R2, T = R1^T (T2 - T1) */
void orc_relative_transform(const double *rec1, int fmt1, const double *rec2, int fmt2, double *R, double *T);
/* CollisionObject::computeAABB for a model under a pose record. */
void orc_object_aabb(const orc_model *m, const double *rec, int pose_format, double *aabb6);

#ifdef __cplusplus
}
#endif
#endif
