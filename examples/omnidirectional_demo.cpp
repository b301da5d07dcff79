// omnidirectional_demo.cpp -- the reference's wiring (main.cpp:38-76: Omnidirectional agent + Map3D workspace)
// on top of the drop-in headers, with the one call-site change the port asks for: edges are handed over in
// batches (Map3D::safeEdges) instead of one safeEdge per planner iteration (planners/rrt.hpp:50).
//
//   omnidirectional_demo <agent.mptm> <environment.mptm> <n_edges> <out.bin>
//
// Writes, for the tests: n_edges (u64), then per edge start xyz, end xyz (6 f64), the batched verdict (u8)
// and the one-at-a-time verdict of Map3D::safeEdge (u8).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mpt_b200/map3d.hpp"

// Minimal agent with the members Map3D needs (Edge, getMeshes, getPoses) and the reference's edge
// interpolation rule (agents/omnidirectional.hpp:202-247).
class Omnidirectional {
public:
	struct State {
		double v[3];
		double x() const { return v[0]; }
		double y() const { return v[1]; }
		double z() const { return v[2]; }
	};
	struct Edge {
		State start, end;
		double cost;
	};
	explicit Omnidirectional(const InstanceFileMap &args) : mesh(args.value("Agent Mesh")) {}

	std::vector<const SimpleAgentMeshHandler *> getMeshes() const {
		return std::vector<const SimpleAgentMeshHandler *>(1, &mesh);
	}

	std::vector<std::vector<fcl::Transform3f> > getPoses(const Edge &edge, double dt) const {
		std::vector<std::vector<fcl::Transform3f> > ret;
		const double sx = edge.start.x(), sy = edge.start.y(), sz = edge.start.z();
		const double dx = edge.end.x() - sx, dy = edge.end.y() - sy, dz = edge.end.z() - sz;
		const double dist = sqrt(dx * dx + dy * dy + dz * dz);
		const unsigned int iterations = dist / dt;
		auto push = [&ret](double x, double y, double z) {
			ret.emplace_back();
			ret.back().push_back(fcl::Transform3f(fcl::Vec3f(x, y, z)));
		};
		if (iterations < 1) {
			push(sx, sy, sz);
			push(edge.end.x(), edge.end.y(), edge.end.z());
		} else {
			const double step = dt / dist;
			for (unsigned int i = 0; i < iterations; ++i) {
				const double s = step * (double)i;
				push(sx + s * dx, sy + s * dy, sz + s * dz);
			}
			if ((double)iterations * dt < dist) push(edge.end.x(), edge.end.y(), edge.end.z());
		}
		return ret;
	}

private:
	SimpleAgentMeshHandler mesh;
};

static uint64_t sm_state = 0x5EED0001ull;
static double uniform01() {  // splitmix64
	uint64_t z = (sm_state += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z ^= z >> 31;
	return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char **argv) {
	if (argc < 5) {
		fprintf(stderr, "usage: %s <agent.mptm> <environment.mptm> <n_edges> <out.bin>\n", argv[0]);
		return 2;
	}
	InstanceFileMap args;  // the keys of omnidirectional.inst
	args.set("Agent Mesh", argv[1]);
	args.set("Environment Mesh", argv[2]);
	args.set("Environment Location", "0 0 0 1 0 0 0");
	args.set("Environment Bounding Box", "-10 10 -10 10 -10 10");
	const double ccDT = 0.1;  // "Collision Check Delta t"
	const size_t n = (size_t)atol(argv[3]);

	Omnidirectional agent(args);
	Map3D<Omnidirectional> workspace(args);

	std::vector<Omnidirectional::Edge> edges;
	for (size_t i = 0; i < n; ++i) {
		Omnidirectional::Edge e;
		for (int k = 0; k < 3; ++k) e.start.v[k] = -3.0 + 6.0 * uniform01();
		double d[3], len = 0;
		for (int k = 0; k < 3; ++k) {
			d[k] = 2.0 * uniform01() - 1.0;
			len += d[k] * d[k];
		}
		len = sqrt(len);
		for (int k = 0; k < 3; ++k) e.end.v[k] = e.start.v[k] + d[k] / len;  // unit random step, as randomSteer
		e.cost = len;
		edges.push_back(e);
	}
	const std::vector<bool> batched = workspace.safeEdges(agent, edges, ccDT);
	// the same edges with the interpolation done on the device
	std::vector<double> startEnd;
	for (const auto &e : edges) {
		startEnd.insert(startEnd.end(), e.start.v, e.start.v + 3);
		startEnd.insert(startEnd.end(), e.end.v, e.end.v + 3);
	}
	const std::vector<bool> onDevice = workspace.safeStraightEdges(agent, startEnd, ccDT);
	// and in three slices that are all in flight before the first result is read (Map3D::submitEdges)
	std::vector<Map3D<Omnidirectional>::PendingEdges> inFlight;
	const size_t cut[4] = {0, n / 3, n / 2, n};
	for (int k = 0; k < 3; ++k)
		inFlight.push_back(workspace.submitEdges(
		    agent, std::vector<Omnidirectional::Edge>(edges.begin() + (long)cut[k], edges.begin() + (long)cut[k + 1]), ccDT));
	std::vector<bool> sliced;
	for (int k = 2; k >= 0; --k) {  // collected newest first: a wait also completes the older tickets
		const std::vector<bool> part = inFlight[(size_t)k].get();
		sliced.insert(sliced.begin(), part.begin(), part.end());
	}
	FILE *f = fopen(argv[4], "wb");
	const uint64_t n64 = n;
	fwrite(&n64, 8, 1, f);
	size_t n_safe = 0, n_diff = 0;
	for (size_t i = 0; i < n; ++i) {
		const bool single = workspace.safeEdge(agent, edges[i], ccDT);
		const uint8_t b = batched[i], s = single;
		n_safe += b;
		n_diff += b != s;
		n_diff += onDevice[i] != batched[i];
		n_diff += sliced[i] != batched[i];
		fwrite(edges[i].start.v, 8, 3, f);
		fwrite(edges[i].end.v, 8, 3, f);
		fwrite(&b, 1, 1, f);
		fwrite(&s, 1, 1, f);
	}
	fclose(f);
	printf("edges %zu safe %zu batched!=single %zu\n", n, n_safe, n_diff);
	return n_diff == 0 ? 0 : 1;
}
