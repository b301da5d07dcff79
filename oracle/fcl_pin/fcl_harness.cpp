// fcl_harness.cpp -- pins the oracle (and with it every parity claim of this repository) against the REAL reference:
// it includes /root/reference/utilities/meshhandler.hpp UNMODIFIED and calls MeshHandler::isInCollision
// (meshhandler.hpp:187-243) once per pose-set, exactly as Map3D::safePoses does, on top of a real FCL 0.3.2 + Assimp
// + Boost installation (the versions /root/reference/README.md:13 names).  None of those exist in the build
// container of this repository, so this file cannot be compiled there: it is the one-command route for anyone who
// has FCL ("make -C oracle _ref && python tools/pin_golden_with_fcl.py").  TEST INFRASTRUCTURE, never shipped.
//
//   fcl_harness <env mesh file> "<x y z qw qx qy qz>" <agent mesh file> <poses.bin> <out.bits> [--self]
//   fcl_harness --dump-meshes <mesh file>            (what AssimpMeshLoader::get hands over, for mesh-ingestion parity)
//
// poses.bin : "MPTPIN01", u32 format (0 = RT12: 12 doubles R row-major + T -> fcl::Transform3f(Matrix3f, Vec3f);
//             1 = TQ7: x y z qw qx qy qz -> fcl::Transform3f(Quaternion3f(w,x,y,z), Vec3f)), u32 n_links,
//             u64 n_pose_sets, then n_pose_sets * n_links records of fp64.
// out.bits  : ceil(n/32) u32 words, bit i%32 of word i/32 = isInCollision(pose-set i)  (the layout of
//             mptcc_check_poses and of tests/golden/*.npz "bits").
// Every link uses the same SimpleAgentMeshHandler, as SnakeTrailers::getMeshes does (snake_trailers.hpp:215-218).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "utilities/meshhandler.hpp"  // the reference's own header, from -I<reference root>

static bool read_exact(FILE *f, void *p, size_t n) { return fread(p, 1, n, f) == n; }

static int dump_meshes(const char *file) {
	AssimpMeshLoader loader(file);
	if (loader.error) return 2;
	std::vector<std::vector<fcl::Vec3f> > vertices;
	std::vector<std::vector<fcl::Triangle> > triangles;
	std::vector<std::vector<double> > normals;
	loader.get(vertices, triangles, normals);
	printf("%zu\n", vertices.size());
	for (size_t i = 0; i < vertices.size(); ++i) {
		// FNV-1a over the float64 coordinates and the indices, the digest tools/pin_golden_with_fcl.py recomputes
		uint64_t h = 1469598103934665603ull;
		auto mix = [&h](const void *p, size_t n) {
			const unsigned char *b = static_cast<const unsigned char *>(p);
			for (size_t k = 0; k < n; ++k) {
				h ^= b[k];
				h *= 1099511628211ull;
			}
		};
		for (const fcl::Vec3f &v : vertices[i])
			for (int k = 0; k < 3; ++k) {
				const double d = v[k];
				mix(&d, 8);
			}
		for (const fcl::Triangle &t : triangles[i])
			for (int k = 0; k < 3; ++k) {
				const uint32_t idx = (uint32_t)t[k];
				mix(&idx, 4);
			}
		printf("%zu %zu %016llx\n", vertices[i].size(), triangles[i].size(), (unsigned long long)h);
	}
	return 0;
}

int main(int argc, char **argv) {
	if (argc == 3 && std::strcmp(argv[1], "--dump-meshes") == 0) return dump_meshes(argv[2]);
	if (argc < 6) {
		fprintf(stderr, "usage: %s <env mesh> \"<x y z qw qx qy qz>\" <agent mesh> <poses.bin> <out.bits> [--self]\n", argv[0]);
		return 64;
	}
	const bool self = argc > 6 && std::strcmp(argv[6], "--self") == 0;
	StaticEnvironmentMeshHandler environment(argv[1], argv[2]);  // meshhandler.hpp:18-55 (parseTransform inside)
	SimpleAgentMeshHandler agentHandler(argv[3]);                // meshhandler.hpp:114-135 (last sub-mesh wins)

	FILE *in = fopen(argv[4], "rb");
	if (!in) return 66;
	char magic[8];
	uint32_t fmt = 0, n_links = 0;
	uint64_t n = 0;
	if (!read_exact(in, magic, 8) || std::memcmp(magic, "MPTPIN01", 8) != 0 || !read_exact(in, &fmt, 4) ||
	    !read_exact(in, &n_links, 4) || !read_exact(in, &n, 8) || fmt > 1 || n_links < 1)
		return 65;
	const size_t rec = fmt == 1 ? 7 : 12;
	std::vector<double> buf(rec * n_links);
	std::vector<uint32_t> bits((n + 31) / 32, 0u);
	std::vector<const SimpleAgentMeshHandler *> agent(n_links, &agentHandler);
	for (uint64_t i = 0; i < n; ++i) {
		if (!read_exact(in, buf.data(), sizeof(double) * buf.size())) return 65;
		std::vector<std::vector<fcl::Transform3f> > poses(1);
		for (uint32_t l = 0; l < n_links; ++l) {
			const double *r = &buf[rec * l];
			if (fmt == 1)
				poses[0].push_back(fcl::Transform3f(fcl::Quaternion3f(r[3], r[4], r[5], r[6]), fcl::Vec3f(r[0], r[1], r[2])));
			else
				poses[0].push_back(fcl::Transform3f(fcl::Matrix3f(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8]),
				                                    fcl::Vec3f(r[9], r[10], r[11])));
		}
		if (MeshHandler::isInCollision(environment, agent, poses, self)) bits[i >> 5] |= 1u << (i & 31u);
	}
	fclose(in);
	FILE *out = fopen(argv[5], "wb");
	if (!out) return 73;
	fwrite(bits.data(), 4, bits.size(), out);
	fclose(out);
	return 0;
}
