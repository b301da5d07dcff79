// meshhandler.hpp -- drop-in for /root/reference/utilities/meshhandler.hpp on top of libmptcc.
//
// Same class and method names, argument meaning and error behaviour as the reference:
//   StaticEnvironmentMeshHandler(file, pose)   meshhandler.hpp:18-55   (exit(0) when the mesh cannot be loaded)
//   SimpleAgentMeshHandler(file)               meshhandler.hpp:114-135 (keeps the LAST non-empty sub-mesh)
//   MeshHandler::isInCollision(env, agent, poses, checkSelfCollision, debug)   meshhandler.hpp:187-243
// plus the batched form the call sites switch to:
//   MeshHandler::collidingPoseSets(env, agent, poses) -> one bool per pose-set (one isInCollision each).
//
// Mesh files: Assimp is replaced by mesh_loaders.hpp (.dae <triangles>, .3ds, and the flat .mptm format) with
// the sub-mesh semantics of AssimpMeshLoader::get (/root/reference/utilities/assimp_mesh_loader.hpp:20-60).
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "../mptcc.h"
#include "fcl_types.hpp"
#include "mesh_loaders.hpp"

namespace fcl_helpers {
// fcl_helpers::parseTransform (/root/reference/utilities/fcl_helpers.hpp:16-25): "x y z qw qx qy qz", read
// left to right (the reference's unsequenced *++token arguments only work that way).
inline fcl::Transform3f parseTransform(const std::string &transform) {
	std::istringstream in(transform);
	double v[7] = {0, 0, 0, 1, 0, 0, 0};
	for (int i = 0; i < 7; ++i) in >> v[i];
	return fcl::Transform3f(fcl::Quaternion3f(v[3], v[4], v[5], v[6]), fcl::Vec3f(v[0], v[1], v[2]));
}
}  // namespace fcl_helpers

namespace mpt_b200 {

inline mptcc_mesh as_c_mesh(const SubMesh &m) {
	mptcc_mesh c;
	c.xyz = m.xyz.data();
	c.nv = (uint32_t)(m.xyz.size() / 3);
	c.tri = m.tri.data();
	c.nt = (uint32_t)(m.tri.size() / 3);
	return c;
}

}  // namespace mpt_b200

class SimpleAgentMeshHandler {
public:
	// mergeSubMeshes = false reproduces the reference: agentModel is overwritten for every non-empty
	// sub-mesh, so only the last one collides (meshhandler.hpp:124-134).
	explicit SimpleAgentMeshHandler(const std::string &filename, bool mergeSubMeshes = false) : id(nextId()) {
		mpt_b200::MeshFileLoader loader(filename);
		if (loader.error) exit(0);
		for (const auto &m : loader.meshes) {
			if (m.xyz.empty() || m.tri.empty()) continue;
			if (!mergeSubMeshes) {
				model = m;
			} else {
				const uint32_t base = (uint32_t)(model.xyz.size() / 3);
				model.xyz.insert(model.xyz.end(), m.xyz.begin(), m.xyz.end());
				for (uint32_t i : m.tri) model.tri.push_back(i + base);
			}
		}
	}
	const mpt_b200::SubMesh &getModel() const { return model; }
	// Identity of this handler's model for the environment's "agent currently installed" cache: unique per
	// constructed handler, so a handler that is freed and another one allocated at the same address never aliases.
	unsigned long long getId() const { return id; }

private:
	static unsigned long long nextId() {
		static unsigned long long counter = 0;
		return ++counter;
	}
	mpt_b200::SubMesh model;
	unsigned long long id;
};

class StaticEnvironmentMeshHandler {
public:
	StaticEnvironmentMeshHandler(const std::string &filename, const std::string &pose, int nDevices = 1)
	    : ctx(nullptr), agentKey() {
		mpt_b200::MeshFileLoader loader(filename);
		if (loader.error) exit(0);
		const fcl::Transform3f tf = fcl_helpers::parseTransform(pose);
		for (const auto &m : loader.meshes)
			if (!m.xyz.empty() && !m.tri.empty()) objects.push_back(m);  // one CollisionObject per non-empty sub-mesh
		const int rc = mptcc_create(&ctx, nDevices, nullptr);
		if (rc != MPTCC_OK) {
			fprintf(stderr, "mptcc_create failed (%d): no usable B200 -- there is no CPU fallback\n", rc);
			exit(1);
		}
		std::vector<mptcc_mesh> cm;
		std::vector<double> pose7;
		for (const auto &m : objects) {
			cm.push_back(mpt_b200::as_c_mesh(m));
			const double p[7] = {tf.getTranslation()[0], tf.getTranslation()[1], tf.getTranslation()[2],
			                     tf.getQuatRotation().getW(), tf.getQuatRotation().getX(), tf.getQuatRotation().getY(),
			                     tf.getQuatRotation().getZ()};
			pose7.insert(pose7.end(), p, p + 7);
		}
		check(mptcc_set_environment(ctx, (int)cm.size(), cm.data(), pose7.data()), "mptcc_set_environment");
	}
	~StaticEnvironmentMeshHandler() { mptcc_destroy(ctx); }
	StaticEnvironmentMeshHandler(const StaticEnvironmentMeshHandler &) = delete;
	StaticEnvironmentMeshHandler &operator=(const StaticEnvironmentMeshHandler &) = delete;

	// The context bound to this environment, with `agent` installed as the current link models.
	// (mptcc_set_agent first waits for every outstanding submit of this context, so replacing the agent while
	// PendingVerdicts / PendingEdges are in flight is safe.)
	mptcc_ctx *getCollisionChecker(const std::vector<const SimpleAgentMeshHandler *> &agent) const {
		std::vector<unsigned long long> key;
		for (const auto *a : agent) key.push_back(a->getId());
		if (key != agentKey) {
			std::vector<mptcc_mesh> cm;
			for (const auto *a : agent) cm.push_back(mpt_b200::as_c_mesh(a->getModel()));
			check(mptcc_set_agent(ctx, (int)cm.size(), cm.data()), "mptcc_set_agent");
			agentKey = key;
		}
		return ctx;
	}
	void check(int rc, const char *what) const {
		if (rc != MPTCC_OK) {
			fprintf(stderr, "%s failed (%d): %s\n", what, rc, mptcc_last_error(ctx));
			exit(1);
		}
	}

private:
	std::vector<mpt_b200::SubMesh> objects;
	mptcc_ctx *ctx;
	mutable std::vector<unsigned long long> agentKey;
};

namespace mpt_b200 {
// Pinned staging blocks for the asynchronous calls, reused between batches (cudaMallocHost is far too slow
// to sit on a per-call path).  Not thread-safe, like the rest of the shim.
class PinnedPool {
public:
	static void *take(size_t bytes) {
		std::vector<std::pair<void *, size_t> > &f = blocks();
		size_t best = f.size();
		for (size_t i = 0; i < f.size(); ++i)
			if (f[i].second >= bytes && (best == f.size() || f[i].second < f[best].second)) best = i;
		if (best != f.size()) {
			void *p = f[best].first;
			sizes()[p] = f[best].second;
			f.erase(f.begin() + (long)best);
			return p;
		}
		const size_t cap = bytes < 4096 ? 4096 : bytes + bytes / 2;
		void *p = mptcc_alloc_pinned(cap);
		if (!p) {
			fprintf(stderr, "mpt_b200: cannot allocate %zu bytes of pinned memory\n", cap);
			exit(1);
		}
		sizes()[p] = cap;
		return p;
	}
	static void give(void *p) {
		if (p) blocks().push_back(std::make_pair(p, sizes()[p]));
	}

private:
	static std::vector<std::pair<void *, size_t> > &blocks() {
		static std::vector<std::pair<void *, size_t> > b;
		return b;
	}
	static std::map<void *, size_t> &sizes() {
		static std::map<void *, size_t> s;
		return s;
	}
};
}  // namespace mpt_b200

// Verdicts of a batch whose device work is still in flight (MeshHandler::submitPoseSets): get() waits for it.
class PendingVerdicts {
public:
	PendingVerdicts() : env(nullptr), ctx(nullptr), done(true) {}
	~PendingVerdicts() { finish(); }
	// one bool per pose-set of the submitted batch, true = in collision
	const std::vector<bool> &get() {
		finish();
		return verdict;
	}

private:
	friend class MeshHandler;
	PendingVerdicts(const PendingVerdicts &);
	PendingVerdicts &operator=(const PendingVerdicts &);
	struct Part {
		double *rec;
		uint32_t *bits;
		std::vector<size_t> index;
		uint64_t ticket;
	};
	void finish();
	const StaticEnvironmentMeshHandler *env;
	mptcc_ctx *ctx;
	std::vector<Part> parts;
	std::vector<bool> verdict;
	bool done;
};

class MeshHandler {
public:
	// Pose record format of one transform: the most compact libmptcc format that carries exactly what FCL would
	// compute with (fcl_types.hpp): a quaternion-built transform -> TQ7; a matrix-built one -> T3 if the matrix is the
	// identity (Transform3f(Vec3f), Omnidirectional::getPoses), XYZCS if it is the planar rotation the blimp and the
	// snake fill in (R00 = R11 = c, R01 = s, R10 = -s, rest of the identity), RT12 otherwise.
	static int kindOf(const fcl::Transform3f &tf) {
		if (tf.builtFromQuaternion()) return MPTCC_POSE_TQ7;
		const fcl::Matrix3f &R = tf.getRotation();
		if (R(0, 2) == 0 && R(1, 2) == 0 && R(2, 0) == 0 && R(2, 1) == 0 && R(2, 2) == 1 && R(0, 0) == R(1, 1) && R(1, 0) == -R(0, 1))
			return (R(0, 0) == 1 && R(0, 1) == 0) ? MPTCC_POSE_T3 : MPTCC_POSE_XYZCS;
		return MPTCC_POSE_RT12;
	}
	// Format of a whole pose-set: the most general of its links (T3 < XYZCS < RT12); quaternion- and matrix-built
	// links cannot share a record format, no agent of the toolkit produces such a set.
	static int kindOfSet(const std::vector<fcl::Transform3f> &set) {
		static const int rank[4] = {2 /*RT12*/, 3 /*TQ7*/, 0 /*T3*/, 1 /*XYZCS*/};
		int kind = MPTCC_POSE_T3;
		bool any_q = false, any_m = false;
		for (const auto &tf : set) {
			const int k = kindOf(tf);
			(k == MPTCC_POSE_TQ7 ? any_q : any_m) = true;
			if (rank[k] > rank[kind]) kind = k;
		}
		if (any_q && any_m) {
			fprintf(stderr, "MeshHandler: a pose-set mixes quaternion-built and matrix-built transforms\n");
			exit(1);
		}
		return kind;
	}
	static size_t recLen(int kind) { return kind == MPTCC_POSE_TQ7 ? 7 : kind == MPTCC_POSE_T3 ? 3 : kind == MPTCC_POSE_XYZCS ? 5 : 12; }

	// [P][L] transforms -> pose records of format `kind` for the pose-sets listed in `index`.
	static int packPoses(const std::vector<std::vector<fcl::Transform3f> > &poses, const std::vector<size_t> &index,
	                     size_t links, int kind, double *out) {
		for (size_t i : index) {
			const std::vector<fcl::Transform3f> &set = poses[i];
			if (set.size() != links) return -1;
			for (const auto &tf : set) {
				const fcl::Vec3f &T = tf.getTranslation();
				if (kind == MPTCC_POSE_TQ7) {
					const fcl::Quaternion3f &q = tf.getQuatRotation();
					out[0] = T[0];
					out[1] = T[1];
					out[2] = T[2];
					out[3] = q.getW();
					out[4] = q.getX();
					out[5] = q.getY();
					out[6] = q.getZ();
				} else if (kind == MPTCC_POSE_T3 || kind == MPTCC_POSE_XYZCS) {
					out[0] = T[0];
					out[1] = T[1];
					out[2] = T[2];
					if (kind == MPTCC_POSE_XYZCS) {
						out[3] = tf.getRotation()(0, 0);
						out[4] = tf.getRotation()(0, 1);
					}
				} else {
					memcpy(out, tf.getRotation().data(), 9 * sizeof(double));
					out[9] = T[0];
					out[10] = T[1];
					out[11] = T[2];
				}
				out += recLen(kind);
			}
		}
		return 0;
	}
	// pose-set indices per record format; pose-sets without transforms have no agent objects and never collide
	// (the Blimp::getPoses stub, /root/reference/agents/blimp.hpp:219-223)
	static void splitByKind(const std::vector<std::vector<fcl::Transform3f> > &poses, std::vector<size_t> index[4]) {
		for (size_t i = 0; i < poses.size(); ++i)
			if (!poses[i].empty()) index[kindOfSet(poses[i])].push_back(i);
	}

	// One verdict per pose-set (true = in collision): what one MeshHandler::isInCollision call per pose-set
	// would return.  This is the call the batched planners use.
	static std::vector<bool> collidingPoseSets(const StaticEnvironmentMeshHandler &environment,
	                                           const std::vector<const SimpleAgentMeshHandler *> &agent,
	                                           const std::vector<std::vector<fcl::Transform3f> > &poses) {
		// one edge or so: the blocking C call has a copy-free path for up to 256 KB of poses
		if (!poses.empty() && !agent.empty() && poses.size() * agent.size() * 12 * sizeof(double) <= (256u << 10))
			return collidingPoseSetsBlocking(environment, agent, poses);
		return submitPoseSets(environment, agent, poses)->get();
	}

	static std::vector<bool> collidingPoseSetsBlocking(const StaticEnvironmentMeshHandler &environment,
	                                                   const std::vector<const SimpleAgentMeshHandler *> &agent,
	                                                   const std::vector<std::vector<fcl::Transform3f> > &poses) {
		std::vector<bool> verdict(poses.size(), false);
		mptcc_ctx *ctx = environment.getCollisionChecker(agent);
		std::vector<double> rec;
		std::vector<uint32_t> bits;
		std::vector<size_t> byKind[4];
		splitByKind(poses, byKind);
		for (int kind = 0; kind < 4; ++kind) {
			const std::vector<size_t> &index = byKind[kind];
			if (index.empty()) continue;
			rec.resize(index.size() * agent.size() * recLen(kind));
			bits.assign((index.size() + 31) / 32, 0u);
			if (packPoses(poses, index, agent.size(), kind, rec.data()) != 0) {
				fprintf(stderr, "MeshHandler: every pose-set needs one transform per agent mesh\n");
				exit(1);
			}
			environment.check(mptcc_check_poses(ctx, index.size(), rec.data(), kind, bits.data()), "mptcc_check_poses");
			for (size_t k = 0; k < index.size(); ++k) verdict[index[k]] = (bits[k >> 5] >> (k & 31)) & 1u;
		}
		return verdict;
	}

	// collidingPoseSets without the wait: the copies and kernels are enqueued (mptcc_submit_poses, pinned
	// staging) and the caller goes on preparing its next batch; verdicts = result->get().  Results of earlier
	// submits complete first; up to four may be outstanding.
	static std::shared_ptr<PendingVerdicts> submitPoseSets(const StaticEnvironmentMeshHandler &environment,
	                                                       const std::vector<const SimpleAgentMeshHandler *> &agent,
	                                                       const std::vector<std::vector<fcl::Transform3f> > &poses) {
		std::shared_ptr<PendingVerdicts> out(new PendingVerdicts);
		out->verdict.assign(poses.size(), false);
		if (poses.empty() || agent.empty()) return out;
		out->env = &environment;
		out->ctx = environment.getCollisionChecker(agent);
		// one submit per record format present in the batch
		std::vector<size_t> byKind[4];
		splitByKind(poses, byKind);
		for (int kind = 0; kind < 4; ++kind) {
			if (byKind[kind].empty()) continue;
			PendingVerdicts::Part pt;
			pt.index.swap(byKind[kind]);
			const size_t n = pt.index.size(), words = (n + 31) / 32;
			pt.rec = static_cast<double *>(mpt_b200::PinnedPool::take(n * agent.size() * recLen(kind) * sizeof(double)));
			pt.bits = static_cast<uint32_t *>(mpt_b200::PinnedPool::take(words * sizeof(uint32_t)));
			if (packPoses(poses, pt.index, agent.size(), kind, pt.rec) != 0) {
				fprintf(stderr, "MeshHandler: every pose-set needs one transform per agent mesh\n");
				exit(1);
			}
			pt.ticket = 0;
			out->done = false;
			out->parts.push_back(pt);
			environment.check(mptcc_submit_poses(out->ctx, n, pt.rec, kind, pt.bits, &out->parts.back().ticket), "mptcc_submit_poses");
		}
		return out;
	}

	// straight, translation-only edges (start xyz, end xyz) interpolated ON THE DEVICE with the
	// rule of Omnidirectional::getPoses (agents/omnidirectional.hpp:202-247); result[i] = edge i is safe.
	static std::vector<bool> safeStraightEdges(const StaticEnvironmentMeshHandler &environment,
	                                           const std::vector<const SimpleAgentMeshHandler *> &agent,
	                                           const std::vector<double> &startEnd /* 6 per edge */, double dt) {
		const size_t n = startEnd.size() / 6;
		std::vector<bool> safe(n, true);
		if (n == 0) return safe;
		mptcc_ctx *ctx = environment.getCollisionChecker(agent);
		std::vector<uint8_t> out(n, 0);
		environment.check(mptcc_check_omni_edges(ctx, n, startEnd.data(), dt, out.data()), "mptcc_check_omni_edges");
		for (size_t i = 0; i < n; ++i) safe[i] = out[i] != 0;
		return safe;
	}

	// The reference's signature and meaning: one bool for the whole batch of pose-sets.
	static bool isInCollision(const StaticEnvironmentMeshHandler &environment,
	                          const std::vector<const SimpleAgentMeshHandler *> &agent,
	                          const std::vector<std::vector<fcl::Transform3f> > &poses, bool checkSelfCollision = false,
	                          bool debug = false) {
		(void)debug;  // contact points are not computed (enable_contact = false in the reference as well)
		// meshhandler.hpp:205-219: the links of every pose-set are also tested against each other
		const bool self = checkSelfCollision && !agent.empty() && !poses.empty();
		mptcc_ctx *ctx = self ? environment.getCollisionChecker(agent) : nullptr;
		if (self) environment.check(mptcc_set_option(ctx, "self_collision", 1), "mptcc_set_option");
		const std::vector<bool> v = collidingPoseSets(environment, agent, poses);
		if (self) environment.check(mptcc_set_option(ctx, "self_collision", 0), "mptcc_set_option");
		for (bool b : v)
			if (b) return true;
		return false;
	}
};

inline void PendingVerdicts::finish() {
	if (done) return;
	done = true;
	for (Part &pt : parts) {
		env->check(mptcc_wait(ctx, pt.ticket), "mptcc_wait");
		for (size_t k = 0; k < pt.index.size(); ++k) verdict[pt.index[k]] = (pt.bits[k >> 5] >> (k & 31)) & 1u;
		mpt_b200::PinnedPool::give(pt.rec);
		mpt_b200::PinnedPool::give(pt.bits);
	}
	parts.clear();
}
