// bvh.hpp -- host-side OBB tree builder of libmptcc (product code, C++17).
//
// Produces, for one (vertices, triangles) sub-model, the same tree fcl::BVHModel<fcl::OBBRSS> builds in
// beginModel/addSubModel/endModel (called from /root/reference/utilities/meshhandler.hpp:44-47,130-133):
// 2n-1 nodes, one triangle per leaf, OBB fitted by covariance + cyclic Jacobi, mean split along the
// major axis, children numbered in FCL's pre-order allocation order.  The fp64 node values have to be
// bit-identical to FCL's because the device replays FCL's OBB tests in fp64 (collide_kernel.cu,
// "path check").  [FCL-recalled] upstream: src/BVH/BVH_model.cpp, BV_fitter.cpp, BV_splitter.cpp,
// BVH_utility.cpp, include/fcl/math/matrix_3f.h (eigen).
#pragma once

#include <cstdint>
#include <vector>

namespace mptcc {

struct HostBvh {
	// inputs (copied)
	std::vector<double> xyz;    // [nv*3]
	std::vector<uint32_t> tri;  // [nt*3]
	// nodes, SoA
	std::vector<double> axis;   // [n*9]  axis[k] = 3 consecutive doubles
	std::vector<double> center; // [n*3]  OBB::To
	std::vector<double> extent; // [n*3]
	std::vector<int32_t> first_child;     // >=0: children (fc, fc+1); <0: leaf, primitive = -(fc+1)
	std::vector<int32_t> first_primitive;
	std::vector<int32_t> num_primitives;
	std::vector<uint32_t> primitive_indices;  // permutation of 0..nt-1
	uint32_t depth = 0;                        // number of levels (a single leaf has depth 1)
	// BVHModel::computeLocalAABB
	double aabb_min[3], aabb_max[3], aabb_center[3], aabb_radius;

	uint32_t num_nodes() const { return (uint32_t)first_child.size(); }
	uint32_t num_tris() const { return (uint32_t)(tri.size() / 3); }
};

// Returns false for an empty model (the reference's handlers skip those, meshhandler.hpp:42,128).
bool build_bvh(const double *xyz, uint32_t nv, const uint32_t *tri, uint32_t nt, HostBvh &out);

// fcl::Transform3f pieces needed on the host ([FCL-recalled] src/math/transform.cpp).
void quat_to_rotation(const double q[4], double R[9]);

}  // namespace mptcc
