// prm_roadmap.cpp -- BASELINE.json config 2: "PRM roadmap build on omnidirectional agent: batched edge validity
// for all k-NN edges", as the call-site change of SURVEY 8f N1.
//
// The reference builds its roadmap one edge at a time (planners/prm/prm.hpp:334-387, addMilestone): for every new
// milestone, its k = 10 nearest earlier milestones, `agent.steer(source, target)` (a full straight edge for the
// Omnidirectional agent), `workspace.safeEdge(agent, edge, collisionCheckDT)`, then add_edge + union_set.
// The validity of an edge never depends on other edges, so a slice of milestones (prm.hpp:222-224: 100 per
// query slice) can collect all of its <= 1000 edges, check them in ONE device call and apply the graph updates
// in the original order.  This program does both and verifies that the two roadmaps are identical.
//
//   prm_roadmap <agent mesh> <environment mesh> <n_milestones> [slice=100]
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>

#include "mpt_b200/map3d.hpp"

struct Agent {  // what Map3D needs from an agent (the reference's Omnidirectional, agents/omnidirectional.hpp)
	struct State { double v[3]; };
	struct Edge { State start, end; double cost; };
	explicit Agent(const std::string &mesh) : mesh(mesh) {}
	std::vector<const SimpleAgentMeshHandler *> getMeshes() const { return std::vector<const SimpleAgentMeshHandler *>(1, &mesh); }
	std::vector<std::vector<fcl::Transform3f> > getPoses(const Edge &e, double dt) const {
		std::vector<std::vector<fcl::Transform3f> > ret;
		const double dx = e.end.v[0] - e.start.v[0], dy = e.end.v[1] - e.start.v[1], dz = e.end.v[2] - e.start.v[2];
		const double dist = sqrt(dx * dx + dy * dy + dz * dz);
		const unsigned int iterations = dist / dt;
		auto push = [&ret](double x, double y, double z) {
			ret.emplace_back();
			ret.back().push_back(fcl::Transform3f(fcl::Vec3f(x, y, z)));
		};
		if (iterations < 1) {
			push(e.start.v[0], e.start.v[1], e.start.v[2]);
			push(e.end.v[0], e.end.v[1], e.end.v[2]);
		} else {
			const double step = dt / dist;
			for (unsigned int i = 0; i < iterations; ++i) {
				const double s = step * (double)i;
				push(e.start.v[0] + s * dx, e.start.v[1] + s * dy, e.start.v[2] + s * dz);
			}
			if ((double)iterations * dt < dist) push(e.end.v[0], e.end.v[1], e.end.v[2]);
		}
		return ret;
	}
	SimpleAgentMeshHandler mesh;
};

struct Roadmap {
	std::vector<std::pair<int, int> > edges;
	std::vector<int> parent;
	int find(int a) { while (parent[a] != a) a = parent[a] = parent[parent[a]]; return a; }
	void add(int a, int b) { edges.push_back(std::make_pair(a, b)); parent[find(a)] = find(b); }
};

static uint64_t sm_state = 0x5EED0002ull;
static double uniform01() {
	uint64_t z = (sm_state += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z ^= z >> 31;
	return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char **argv) {
	if (argc < 4) {
		fprintf(stderr, "usage: %s <agent mesh> <environment mesh> <n_milestones> [slice]\n", argv[0]);
		return 2;
	}
	const int n = atoi(argv[3]), slice = argc > 4 ? atoi(argv[4]) : 100, k = 10;
	const double ccDT = 0.1;
	InstanceFileMap args;
	args.set("Environment Mesh", argv[2]);
	args.set("Environment Location", "0 0 0 1 0 0 0");
	args.set("Environment Bounding Box", "-10 10 -10 10 -10 10");
	Agent agent(argv[1]);
	Map3D<Agent> workspace(args);

	std::vector<Agent::State> pts(n);
	for (auto &p : pts)
		for (int a = 0; a < 3; ++a) p.v[a] = -10.0 + 20.0 * uniform01();
	// k nearest EARLIER milestones of milestone i (brute force; the reference uses a FLANN kd-tree)
	auto neighbours = [&](int i) {
		std::vector<std::pair<double, int> > d;
		for (int j = 0; j < i; ++j) {
			double s = 0;
			for (int a = 0; a < 3; ++a) s += (pts[i].v[a] - pts[j].v[a]) * (pts[i].v[a] - pts[j].v[a]);
			d.push_back(std::make_pair(s, j));
		}
		std::sort(d.begin(), d.end());
		if ((int)d.size() > k) d.resize(k);
		return d;
	};
	auto clock = [] { return std::chrono::steady_clock::now(); };
	auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
		return std::chrono::duration<double, std::milli>(b - a).count();
	};

	// ---- the reference's order: one safeEdge per edge ---------------------------------------------------
	Roadmap seq;
	seq.parent.resize(n);
	std::iota(seq.parent.begin(), seq.parent.end(), 0);
	size_t n_checked = 0;
	auto t0 = clock();
	for (int i = 0; i < n; ++i)
		for (const auto &nb : neighbours(i)) {
			Agent::Edge e{pts[i], pts[nb.second], sqrt(nb.first)};
			++n_checked;
			if (workspace.safeEdge(agent, e, ccDT)) seq.add(i, nb.second);
		}
	auto t1 = clock();

	// ---- batched: all edges of a slice in one device call, graph updates in the original order -----------
	Roadmap bat;
	bat.parent.resize(n);
	std::iota(bat.parent.begin(), bat.parent.end(), 0);
	auto t2 = clock();
	for (int s0 = 0; s0 < n; s0 += slice) {
		std::vector<std::pair<int, int> > cand;
		std::vector<double> startEnd;
		for (int i = s0; i < std::min(n, s0 + slice); ++i)
			for (const auto &nb : neighbours(i)) {
				cand.push_back(std::make_pair(i, nb.second));
				startEnd.insert(startEnd.end(), pts[i].v, pts[i].v + 3);
				startEnd.insert(startEnd.end(), pts[nb.second].v, pts[nb.second].v + 3);
			}
		const std::vector<bool> safe = workspace.safeStraightEdges(agent, startEnd, ccDT);
		for (size_t c = 0; c < cand.size(); ++c)
			if (safe[c]) bat.add(cand[c].first, cand[c].second);
	}
	auto t3 = clock();

	int comp_seq = 0, comp_bat = 0;
	for (int i = 0; i < n; ++i) {
		comp_seq += seq.find(i) == i;
		comp_bat += bat.find(i) == i;
	}
	const bool same = seq.edges == bat.edges;
	printf("milestones %d edges checked %zu kept %zu components %d | one-at-a-time %.1f ms | batched(slice %d) %.1f ms | "
	       "identical roadmaps: %s\n", n, n_checked, seq.edges.size(), comp_seq, ms(t0, t1), slice, ms(t2, t3), same ? "yes" : "NO");
	return same && comp_seq == comp_bat ? 0 : 1;
}
