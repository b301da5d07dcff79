// self_collision_check.cpp -- MeshHandler::isInCollision with checkSelfCollision = true
// (/root/reference/utilities/meshhandler.hpp:187-219) through the drop-in headers: a two-link agent of unit
// cubes far from the environment, so that only the links can collide with each other.
//
//   self_collision_check <unit_box.mptm>      prints one line per case: "<name> <self off> <self on>"
#include <cstdio>
#include <vector>

#include "mpt_b200/meshhandler.hpp"

int main(int argc, char **argv) {
	if (argc < 2) {
		fprintf(stderr, "usage: %s <unit_box.mptm>\n", argv[0]);
		return 2;
	}
	StaticEnvironmentMeshHandler environment(argv[1], "100 0 0 1 0 0 0");
	SimpleAgentMeshHandler link0(argv[1]), link1(argv[1]);
	std::vector<const SimpleAgentMeshHandler *> agent;
	agent.push_back(&link0);
	agent.push_back(&link1);
	struct Case {
		const char *name;
		double dx;
	} cases[] = {{"apart", 1.5}, {"touching", 1.0}, {"overlapping", 0.4}, {"just_apart", 1.0 + 1e-12}};
	int bad = 0;
	for (const Case &c : cases) {
		std::vector<std::vector<fcl::Transform3f> > poses(1);
		poses[0].push_back(fcl::Transform3f(fcl::Vec3f(0, 0, 0)));
		poses[0].push_back(fcl::Transform3f(fcl::Vec3f(c.dx, 0, 0)));
		const bool off = MeshHandler::isInCollision(environment, agent, poses);
		const bool on = MeshHandler::isInCollision(environment, agent, poses, /*checkSelfCollision=*/true);
		const bool again = MeshHandler::isInCollision(environment, agent, poses);  // the switch does not stick
		printf("%s %d %d\n", c.name, (int)off, (int)on);
		bad += off || again || on != (c.dx <= 1.0);
	}
	return bad == 0 ? 0 : 1;
}
