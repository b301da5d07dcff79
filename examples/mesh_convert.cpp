// mesh_convert.cpp -- <in.dae|in.3ds|in.mptm> <out.mptm>: the C++ mesh readers of include/mpt_b200/mesh_loaders.hpp
// (Assimp flags-0 sub-mesh semantics) as a converter; tests compare its output with tools/convert_meshes.py.
#include <cstdio>

#include "mpt_b200/mesh_loaders.hpp"

int main(int argc, char **argv) {
	if (argc < 3) {
		fprintf(stderr, "usage: %s <in.dae|in.3ds|in.mptm> <out.mptm>\n", argv[0]);
		return 2;
	}
	mpt_b200::MeshFileLoader loader(argv[1]);
	if (loader.error) return 1;
	size_t nt = 0;
	for (const auto &m : loader.meshes) nt += m.tri.size() / 3;
	printf("%s: %zu sub-meshes, %zu triangles\n", argv[1], loader.meshes.size(), nt);
	return mpt_b200::write_mptm(argv[2], loader.meshes) ? 0 : 1;
}
