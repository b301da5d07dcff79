// rrt_speculative.cpp -- the RRT loop of the reference (/root/reference/planners/rrt.hpp:42-94) with its ONE
// call-site change: instead of one Map3D::safeEdge per iteration, K iterations are SPECULATED from the current tree
// and validated with one device call (Map3D::safeEdges), and the tree that results is EXACTLY the tree the
// one-edge-at-a-time loop grows from the same random stream.
//
// Why that is possible.  Iteration i of the reference loop does: sample a random state, take the tree node nearest
// to it (UniformSampler::getTreeSample, samplers/uniformsampler.hpp:20-35), randomSteer from that node
// (agents/omnidirectional.hpp:168-184), safeEdge, insert the edge's end if safe.  The only coupling between
// iterations is the tree: iteration j > i sees a different nearest node only if iteration i inserted a node that is
// closer to j's sample.  So the random numbers of K iterations are drawn up front (same order as the sequential
// loop draws them), every iteration is steered from its nearest node in the tree AS IT IS NOW, all K edges are
// checked at once, and they are committed in order until one is found whose nearest node has changed because of a
// commit earlier in the same round; that one and the ones after it are re-speculated in the next round with
// their ORIGINAL random numbers.  Nearest neighbours are exact (linear scan, ties by lower index) in both loops.
//
//   rrt_speculative <agent.mptm> <environment.mptm> <iterations> <K>
// prints the size of both trees, the number of collision calls of each loop and "identical trees: yes|NO".
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "mpt_b200/map3d.hpp"

class Omnidirectional {
public:
	struct State {
		double v[3];
	};
	struct Edge {
		State start, end;
		double cost;
	};
	explicit Omnidirectional(const InstanceFileMap &args) : mesh(args.value("Agent Mesh")) {}
	std::vector<const SimpleAgentMeshHandler *> getMeshes() const { return std::vector<const SimpleAgentMeshHandler *>(1, &mesh); }
	// agents/omnidirectional.hpp:202-247
	std::vector<std::vector<fcl::Transform3f> > getPoses(const Edge &edge, double dt) const {
		std::vector<std::vector<fcl::Transform3f> > ret;
		const double sx = edge.start.v[0], sy = edge.start.v[1], sz = edge.start.v[2];
		const double dx = edge.end.v[0] - sx, dy = edge.end.v[1] - sy, dz = edge.end.v[2] - sz;
		const double dist = sqrt(dx * dx + dy * dy + dz * dz);
		const unsigned int iterations = dist / dt;
		auto push = [&ret](double x, double y, double z) { ret.emplace_back(1, fcl::Transform3f(fcl::Vec3f(x, y, z))); };
		if (iterations < 1) {
			push(sx, sy, sz);
			push(edge.end.v[0], edge.end.v[1], edge.end.v[2]);
		} else {
			const double step = dt / dist;
			for (unsigned int i = 0; i < iterations; ++i) {
				const double s = step * (double)i;
				push(sx + s * dx, sy + s * dy, sz + s * dz);
			}
			if ((double)iterations * dt < dist) push(edge.end.v[0], edge.end.v[1], edge.end.v[2]);
		}
		return ret;
	}
	// randomSteer (agents/omnidirectional.hpp:168-184): a step of length dt in a random direction; the three random
	// numbers are passed in so that both loops can consume the same stream
	Edge randomSteer(const State &start, double dt, const double r[3]) const {
		double d[3] = {2 * r[0] - 1, 2 * r[1] - 1, 2 * r[2] - 1};
		const double n = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
		Edge e;
		e.start = start;
		for (int k = 0; k < 3; ++k) e.end.v[k] = start.v[k] + (n > 0 ? d[k] / n : 0.0) * dt;
		e.cost = dt;
		return e;
	}

private:
	SimpleAgentMeshHandler mesh;
};

struct Draw {
	double sample[3], dir[3];
};

static size_t nearest(const std::vector<Omnidirectional::State> &tree, const double p[3]) {
	size_t best = 0;
	double bd = INFINITY;
	for (size_t i = 0; i < tree.size(); ++i) {
		double d = 0;
		for (int k = 0; k < 3; ++k) d += (tree[i].v[k] - p[k]) * (tree[i].v[k] - p[k]);
		if (d < bd) {
			bd = d;
			best = i;
		}
	}
	return best;
}

int main(int argc, char **argv) {
	if (argc < 5) {
		fprintf(stderr, "usage: %s <agent.mptm> <environment.mptm> <iterations> <K>\n", argv[0]);
		return 2;
	}
	InstanceFileMap args;
	args.set("Agent Mesh", argv[1]);
	args.set("Environment Mesh", argv[2]);
	args.set("Environment Location", "0 0 0 1 0 0 0");
	args.set("Environment Bounding Box", "-4 4 -4 4 -4 4");
	const size_t iterations = (size_t)atol(argv[3]), K = (size_t)atol(argv[4]);
	const double steeringDT = 0.6, collisionCheckDT = 0.1;
	Omnidirectional agent(args);
	Map3D<Omnidirectional> workspace(args);
	const Omnidirectional::State start = {{-3.0, -3.0, -3.0}};

	std::mt19937_64 gen(2024);
	std::uniform_real_distribution<double> u01(0.0, 1.0);
	std::vector<Draw> draws(iterations);
	for (Draw &d : draws) {
		for (int k = 0; k < 3; ++k) d.sample[k] = -4.0 + 8.0 * u01(gen);
		for (int k = 0; k < 3; ++k) d.dir[k] = u01(gen);
	}

	// ---- the reference loop: one safeEdge per iteration ------------------------------------------------------------
	std::vector<Omnidirectional::State> treeSeq(1, start);
	size_t callsSeq = 0;
	for (size_t i = 0; i < iterations; ++i) {
		const Omnidirectional::Edge e = agent.randomSteer(treeSeq[nearest(treeSeq, draws[i].sample)], steeringDT, draws[i].dir);
		++callsSeq;
		if (workspace.safeEdge(agent, e, collisionCheckDT)) treeSeq.push_back(e.end);
	}

	// ---- speculative loop: up to K iterations per device call -------------------------------------------------------
	std::vector<Omnidirectional::State> tree(1, start);
	size_t callsSpec = 0, next = 0, respeculated = 0;
	while (next < iterations) {
		const size_t cnt = std::min(K, iterations - next);
		std::vector<Omnidirectional::Edge> edges(cnt);
		std::vector<size_t> from(cnt);
		for (size_t k = 0; k < cnt; ++k) {
			from[k] = nearest(tree, draws[next + k].sample);
			edges[k] = agent.randomSteer(tree[from[k]], steeringDT, draws[next + k].dir);
		}
		const std::vector<bool> safe = workspace.safeEdges(agent, edges, collisionCheckDT);
		++callsSpec;
		const size_t sizeAtSpeculation = tree.size();
		size_t k = 0;
		for (; k < cnt; ++k) {
			// nodes committed earlier in this round: is one of them nearer to this iteration's sample?
			bool stale = false;
			double bd = 0;
			for (int a = 0; a < 3; ++a) bd += (tree[from[k]].v[a] - draws[next + k].sample[a]) * (tree[from[k]].v[a] - draws[next + k].sample[a]);
			for (size_t m = sizeAtSpeculation; m < tree.size() && !stale; ++m) {
				double d = 0;
				for (int a = 0; a < 3; ++a) d += (tree[m].v[a] - draws[next + k].sample[a]) * (tree[m].v[a] - draws[next + k].sample[a]);
				stale = d < bd;
			}
			if (stale) break;
			if (safe[k]) tree.push_back(edges[k].end);
		}
		if (k < cnt) ++respeculated;
		next += k;  // (k >= 1: the first edge of a round is never stale)
	}

	bool same = tree.size() == treeSeq.size();
	for (size_t i = 0; same && i < tree.size(); ++i)
		for (int k = 0; k < 3; ++k) same = same && tree[i].v[k] == treeSeq[i].v[k];
	printf("iterations %zu, K %zu: sequential tree %zu nodes / %zu collision calls; speculative tree %zu nodes / %zu calls "
	       "(%zu rounds cut short by a stale nearest neighbour)\n",
	       iterations, K, treeSeq.size(), callsSeq, tree.size(), callsSpec, respeculated);
	printf("identical trees: %s\n", same ? "yes" : "NO");
	return same ? 0 : 1;
}
