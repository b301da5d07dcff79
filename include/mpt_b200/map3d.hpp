// map3d.hpp -- drop-in for /root/reference/workspaces/map3d.hpp: the same workspace concept
// (getBounds, safeEdge, safePoses) so that RRT, PRM, KPIECE and AnytimeHybridSearch compile unchanged,
// plus the batched entry points their call sites switch to (safeEdges, safePosesBatch).
#pragma once

#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "instancefilemap.hpp"
#include "meshhandler.hpp"

template <class Agent>
class Map3D {
public:
	typedef std::vector<std::pair<double, double> > WorkspaceBounds;
	typedef typename Agent::Edge Edge;

	// Keys as in the reference (map3d.hpp:15-27): "Environment Mesh", "Environment Location",
	// "Environment Bounding Box"; optional "GPU Count" (default 1).
	explicit Map3D(const InstanceFileMap &args)
	    : mesh(args.value("Environment Mesh"), args.value("Environment Location"),
	           args.exists("GPU Count") ? std::stoi(args.value("GPU Count")) : 1),
	      bounds(3) {
		std::istringstream in(args.value("Environment Bounding Box"));
		for (int k = 0; k < 3; ++k) in >> bounds[k].first >> bounds[k].second;
	}

	const WorkspaceBounds &getBounds() const { return bounds; }
	// the environment handler, for callers of MeshHandler::isInCollision that bring their own pose-sets
	const StaticEnvironmentMeshHandler &getEnvironment() const { return mesh; }

	// map3d.hpp:33-37
	bool safeEdge(const Agent &agent, const Edge &edge, double dt, bool checkSelfCollision = false) const {
		auto agentMeshes = agent.getMeshes();
		auto agentPoses = agent.getPoses(edge, dt);
		return !MeshHandler::isInCollision(mesh, agentMeshes, agentPoses, checkSelfCollision);
	}

	// map3d.hpp:39-42.  The reference declares the flat form but forwards it to the nested one (it never
	// compiles there); both shapes that callers use are provided: flat = one single-link pose-set
	// per element.
	bool safePoses(const Agent &agent, const std::vector<fcl::Transform3f> &poses) const {
		std::vector<std::vector<fcl::Transform3f> > nested;
		for (const auto &p : poses) nested.push_back(std::vector<fcl::Transform3f>(1, p));
		return safePoses(agent, nested);
	}
	bool safePoses(const Agent &agent, const std::vector<std::vector<fcl::Transform3f> > &poses) const {
		auto agentMeshes = agent.getMeshes();
		return !MeshHandler::isInCollision(mesh, agentMeshes, poses);
	}
	bool safePose(const Agent &agent, const fcl::Transform3f &pose) const {
		return safePoses(agent, std::vector<fcl::Transform3f>(1, pose));
	}

	// ---- batched forms: one device call for many edges / pose-sets -------------------------------------
	// result[i] == safeEdge(agent, edges[i], dt)
	std::vector<bool> safeEdges(const Agent &agent, const std::vector<Edge> &edges, double dt) const {
		auto agentMeshes = agent.getMeshes();
		std::vector<std::vector<fcl::Transform3f> > all;
		std::vector<size_t> begin(1, 0);
		for (const auto &e : edges) {
			auto poses = agent.getPoses(e, dt);
			all.insert(all.end(), poses.begin(), poses.end());
			begin.push_back(all.size());
		}
		const std::vector<bool> hit = MeshHandler::collidingPoseSets(mesh, agentMeshes, all);
		std::vector<bool> safe(edges.size(), true);
		for (size_t i = 0; i < edges.size(); ++i)
			for (size_t k = begin[i]; k < begin[i + 1]; ++k)
				if (hit[k]) {
					safe[i] = false;
					break;
				}
		return safe;
	}
	// safeEdges in two halves: submitEdges enqueues the device work and returns, PendingEdges::get() waits and
	// folds the pose-set verdicts into one bool per edge.  Lets a planner expand its next slice of edges while
	// this one is on the GPU.
	class PendingEdges {
	public:
		std::vector<bool> get() {
			const std::vector<bool> &hit = verdicts->get();
			std::vector<bool> safe(begin.size() - 1, true);
			for (size_t i = 0; i + 1 < begin.size(); ++i)
				for (size_t k = begin[i]; k < begin[i + 1]; ++k)
					if (hit[k]) {
						safe[i] = false;
						break;
					}
			return safe;
		}

	private:
		friend class Map3D;
		std::shared_ptr<PendingVerdicts> verdicts;
		std::vector<size_t> begin;
	};
	PendingEdges submitEdges(const Agent &agent, const std::vector<Edge> &edges, double dt) const {
		auto agentMeshes = agent.getMeshes();
		std::vector<std::vector<fcl::Transform3f> > all;
		PendingEdges out;
		out.begin.assign(1, 0);
		for (const auto &e : edges) {
			auto poses = agent.getPoses(e, dt);
			all.insert(all.end(), poses.begin(), poses.end());
			out.begin.push_back(all.size());
		}
		out.verdicts = MeshHandler::submitPoseSets(mesh, agentMeshes, all);
		return out;
	}
	// Edges of translation-only agents (Omnidirectional): only (start, end) cross the PCIe bus, the poses
	// are interpolated on the device.  result[i] == safeEdge(agent, Edge(start_i, end_i), dt).
	std::vector<bool> safeStraightEdges(const Agent &agent, const std::vector<double> &startEnd, double dt) const {
		return MeshHandler::safeStraightEdges(mesh, agent.getMeshes(), startEnd, dt);
	}
	// result[i] == safePoses(agent, {poseSets[i]})
	std::vector<bool> safePosesBatch(const Agent &agent, const std::vector<std::vector<fcl::Transform3f> > &poseSets) const {
		auto agentMeshes = agent.getMeshes();
		std::vector<bool> hit = MeshHandler::collidingPoseSets(mesh, agentMeshes, poseSets);
		hit.flip();
		return hit;
	}

private:
	StaticEnvironmentMeshHandler mesh;
	WorkspaceBounds bounds;
};
