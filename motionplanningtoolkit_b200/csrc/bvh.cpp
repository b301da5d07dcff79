// bvh.cpp -- see bvh.hpp.  Build with -ffp-contract=off: every expression must round exactly like
// FCL's x86-64 build (no FMA).
#include "bvh.hpp"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <utility>

namespace mptcc {
namespace {

struct V3 {
	double x, y, z;
};

inline double dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

struct Sym3 {  // full 3x3, row-major
	double m[3][3];
};

struct TriView {
	const double *xyz;
	const uint32_t *tri;
	V3 vertex(uint32_t t, int corner) const {
		const double *p = xyz + 3 * (size_t)tri[3 * (size_t)t + corner];
		return V3{p[0], p[1], p[2]};
	}
};

// [FCL-recalled] getCovariance, BVH_utility.cpp: second moments of the 3n triangle corners.
Sym3 corner_covariance(const TriView &tv, const uint32_t *idx, int n) {
	double s1x = 0, s1y = 0, s1z = 0;
	double sxx = 0, syy = 0, szz = 0, sxy = 0, sxz = 0, syz = 0;
	for (int i = 0; i < n; ++i) {
		const V3 a = tv.vertex(idx[i], 0), b = tv.vertex(idx[i], 1), c = tv.vertex(idx[i], 2);
		s1x += (a.x + b.x + c.x);
		s1y += (a.y + b.y + c.y);
		s1z += (a.z + b.z + c.z);
		sxx += (a.x * a.x + b.x * b.x + c.x * c.x);
		syy += (a.y * a.y + b.y * b.y + c.y * c.y);
		szz += (a.z * a.z + b.z * b.z + c.z * c.z);
		sxy += (a.x * a.y + b.x * b.y + c.x * c.y);
		sxz += (a.x * a.z + b.x * b.z + c.x * c.z);
		syz += (a.y * a.z + b.y * b.z + c.y * c.z);
	}
	const int np = 3 * n;
	Sym3 M;
	M.m[0][0] = sxx - s1x * s1x / np;
	M.m[1][1] = syy - s1y * s1y / np;
	M.m[2][2] = szz - s1z * s1z / np;
	M.m[0][1] = sxy - s1x * s1y / np;
	M.m[1][2] = syz - s1y * s1z / np;
	M.m[0][2] = sxz - s1x * s1z / np;
	M.m[1][0] = M.m[0][1];
	M.m[2][0] = M.m[0][2];
	M.m[2][1] = M.m[1][2];
	return M;
}

// One Jacobi plane rotation applied to the pair (g, h): g' = g - s(h + g tau), h' = h + s(g - h tau).
inline void rot(double &g, double &h, double s, double tau) {
	const double g0 = g, h0 = h;
	g = g0 - s * (h0 + g0 * tau);
	h = h0 + s * (g0 - h0 * tau);
}

// [FCL-recalled] eigen(), matrix_3f.h: Numerical Recipes cyclic Jacobi, at most 50 sweeps.
// evec[k] is ROW k of the accumulated rotation; eval the diagonal.
void jacobi3(const Sym3 &in, double eval[3], double evec[3][3]) {
	double a[3][3];
	std::memcpy(a, in.m, sizeof a);
	double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
	double b[3], d[3], z[3];
	for (int p = 0; p < 3; ++p) {
		b[p] = d[p] = a[p][p];
		z[p] = 0;
	}
	for (int sweep = 0; sweep < 50; ++sweep) {
		double sm = 0;
		for (int p = 0; p < 3; ++p)
			for (int q = p + 1; q < 3; ++q) sm += std::fabs(a[p][q]);
		if (sm == 0.0) {
			for (int k = 0; k < 3; ++k) {
				evec[k][0] = v[k][0];
				evec[k][1] = v[k][1];
				evec[k][2] = v[k][2];
				eval[k] = d[k];
			}
			return;
		}
		const double tresh = (sweep < 3) ? 0.2 * sm / (3 * 3) : 0.0;
		for (int p = 0; p < 3; ++p) {
			for (int q = p + 1; q < 3; ++q) {
				double g = 100.0 * std::fabs(a[p][q]);
				if (sweep > 3 && std::fabs(d[p]) + g == std::fabs(d[p]) && std::fabs(d[q]) + g == std::fabs(d[q])) {
					a[p][q] = 0.0;
				} else if (std::fabs(a[p][q]) > tresh) {
					double h = d[q] - d[p];
					double t;
					if (std::fabs(h) + g == std::fabs(h)) {
						t = (a[p][q]) / h;
					} else {
						const double theta = 0.5 * h / (a[p][q]);
						t = 1.0 / (std::fabs(theta) + std::sqrt(1.0 + theta * theta));
						if (theta < 0.0) t = -t;
					}
					const double c = 1.0 / std::sqrt(1 + t * t);
					const double s = t * c;
					const double tau = s / (1.0 + c);
					h = t * a[p][q];
					z[p] -= h;
					z[q] += h;
					d[p] -= h;
					d[q] += h;
					a[p][q] = 0.0;
					for (int j = 0; j < p; ++j) rot(a[j][p], a[j][q], s, tau);
					for (int j = p + 1; j < q; ++j) rot(a[p][j], a[j][q], s, tau);
					for (int j = q + 1; j < 3; ++j) rot(a[p][j], a[q][j], s, tau);
					for (int j = 0; j < 3; ++j) rot(v[j][p], v[j][q], s, tau);
				}
			}
		}
		for (int p = 0; p < 3; ++p) {
			b[p] += z[p];
			d[p] = b[p];
			z[p] = 0.0;
		}
	}
	// no convergence: upstream leaves the eigenvectors default-constructed (zero)
	for (int k = 0; k < 3; ++k) {
		evec[k][0] = evec[k][1] = evec[k][2] = 0.0;
		eval[k] = 0.0;
	}
}

struct Obb {
	V3 ax[3];
	V3 c;
	double e[3];
};

// [FCL-recalled] BVFitter<OBBRSS>::fit (OBB part) + getExtentAndCenter_mesh.
Obb fit_obb(const TriView &tv, const uint32_t *idx, int n) {
	double s[3], E[3][3];
	jacobi3(corner_covariance(tv, idx, n), s, E);
	int lo, mid, hi;
	if (s[0] > s[1]) {
		hi = 0;
		lo = 1;
	} else {
		lo = 0;
		hi = 1;
	}
	if (s[2] < s[lo]) {
		mid = lo;
		lo = 2;
	} else if (s[2] > s[hi]) {
		mid = hi;
		hi = 2;
	} else {
		mid = 2;
	}
	Obb o;
	o.ax[0] = V3{E[0][hi], E[1][hi], E[2][hi]};
	o.ax[1] = V3{E[0][mid], E[1][mid], E[2][mid]};
	o.ax[2] = V3{E[1][hi] * E[2][mid] - E[1][mid] * E[2][hi], E[0][mid] * E[2][hi] - E[0][hi] * E[2][mid],
	             E[0][hi] * E[1][mid] - E[0][mid] * E[1][hi]};
	double lo_c[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi_c[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
	for (int i = 0; i < n; ++i) {
		for (int corner = 0; corner < 3; ++corner) {
			const V3 p = tv.vertex(idx[i], corner);
			const double proj[3] = {dot(o.ax[0], p), dot(o.ax[1], p), dot(o.ax[2], p)};
			for (int k = 0; k < 3; ++k) {
				if (proj[k] > hi_c[k]) hi_c[k] = proj[k];
				if (proj[k] < lo_c[k]) lo_c[k] = proj[k];
			}
		}
	}
	const double m0 = (hi_c[0] + lo_c[0]) / 2, m1 = (hi_c[1] + lo_c[1]) / 2, m2 = (hi_c[2] + lo_c[2]) / 2;
	o.c = V3{o.ax[0].x * m0 + o.ax[1].x * m1 + o.ax[2].x * m2, o.ax[0].y * m0 + o.ax[1].y * m1 + o.ax[2].y * m2,
	         o.ax[0].z * m0 + o.ax[1].z * m1 + o.ax[2].z * m2};
	o.e[0] = (hi_c[0] - lo_c[0]) / 2;
	o.e[1] = (hi_c[1] - lo_c[1]) / 2;
	o.e[2] = (hi_c[2] - lo_c[2]) / 2;
	return o;
}

struct Task {
	int node, first, count, level;
};

}  // namespace

bool build_bvh(const double *xyz, uint32_t nv, const uint32_t *tri, uint32_t nt, HostBvh &out) {
	if (nv == 0 || nt == 0 || !xyz || !tri) return false;
	for (size_t i = 0; i < 3 * (size_t)nt; ++i)
		if (tri[i] >= nv) return false;  // a face index beyond the vertex array (malformed file / bad mptcc_mesh)
	out.xyz.assign(xyz, xyz + 3 * (size_t)nv);
	out.tri.assign(tri, tri + 3 * (size_t)nt);
	const size_t nn = 2 * (size_t)nt - 1;
	out.axis.assign(nn * 9, 0.0);
	out.center.assign(nn * 3, 0.0);
	out.extent.assign(nn * 3, 0.0);
	out.first_child.assign(nn, 0);
	out.first_primitive.assign(nn, 0);
	out.num_primitives.assign(nn, 0);
	out.primitive_indices.resize(nt);
	for (uint32_t i = 0; i < nt; ++i) out.primitive_indices[i] = i;
	out.depth = 0;

	const TriView tv{out.xyz.data(), out.tri.data()};
	// FCL recurses (node, left subtree, right subtree) and hands out child ids from a running counter
	// when a node is split; a LIFO work list with the right child pushed first visits nodes in the
	// same order, hence allocates the same ids.
	std::vector<Task> work;
	work.push_back(Task{0, 0, (int)nt, 1});
	int next_id = 1;
	while (!work.empty()) {
		const Task t = work.back();
		work.pop_back();
		uint32_t *idx = out.primitive_indices.data() + t.first;
		const Obb o = fit_obb(tv, idx, t.count);
		double *A = &out.axis[9 * (size_t)t.node];
		for (int k = 0; k < 3; ++k) {
			A[3 * k + 0] = o.ax[k].x;
			A[3 * k + 1] = o.ax[k].y;
			A[3 * k + 2] = o.ax[k].z;
			out.extent[3 * (size_t)t.node + k] = o.e[k];
		}
		out.center[3 * (size_t)t.node + 0] = o.c.x;
		out.center[3 * (size_t)t.node + 1] = o.c.y;
		out.center[3 * (size_t)t.node + 2] = o.c.z;
		out.first_primitive[t.node] = t.first;
		out.num_primitives[t.node] = t.count;
		if ((uint32_t)t.level > out.depth) out.depth = (uint32_t)t.level;
		if (t.count == 1) {
			out.first_child[t.node] = -((int32_t)idx[0] + 1);
			continue;
		}
		// [FCL-recalled] BVSplitter<OBBRSS>::computeRule_mean: split along axis[0] at the mean corner.
		const V3 sv = o.ax[0];
		double cx = 0.0, cy = 0.0, cz = 0.0;
		for (int i = 0; i < t.count; ++i) {
			const V3 a = tv.vertex(idx[i], 0), b = tv.vertex(idx[i], 1), c = tv.vertex(idx[i], 2);
			cx += (a.x + b.x + c.x);
			cy += (a.y + b.y + c.y);
			cz += (a.z + b.z + c.z);
		}
		const double split_value = (cx * sv.x + cy * sv.y + cz * sv.z) / (3 * t.count);
		const int fc = next_id;
		next_id += 2;
		out.first_child[t.node] = fc;
		int left = 0;
		for (int i = 0; i < t.count; ++i) {
			const V3 a = tv.vertex(idx[i], 0), b = tv.vertex(idx[i], 1), c = tv.vertex(idx[i], 2);
			const V3 centroid{(a.x + b.x + c.x) / 3.0, (a.y + b.y + c.y) / 3.0, (a.z + b.z + c.z) / 3.0};
			if (!(dot(sv, centroid) > split_value)) {  // not on the right side: move to the front
				std::swap(idx[i], idx[left]);
				++left;
			}
		}
		if (left == 0 || left == t.count) left = t.count / 2;
		work.push_back(Task{fc + 1, t.first + left, t.count - left, t.level + 1});
		work.push_back(Task{fc, t.first, left, t.level + 1});
	}

	// [FCL-recalled] BVHModel::computeLocalAABB over ALL model vertices.
	double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
	for (uint32_t i = 0; i < nv; ++i)
		for (int k = 0; k < 3; ++k) {
			const double c = out.xyz[3 * (size_t)i + k];
			if (c < mn[k]) mn[k] = c;
			if (c > mx[k]) mx[k] = c;
		}
	double r2 = 0;
	for (int k = 0; k < 3; ++k) {
		out.aabb_min[k] = mn[k];
		out.aabb_max[k] = mx[k];
		out.aabb_center[k] = (mn[k] + mx[k]) * 0.5;
	}
	for (uint32_t i = 0; i < nv; ++i) {
		const double dx = out.aabb_center[0] - out.xyz[3 * (size_t)i + 0];
		const double dy = out.aabb_center[1] - out.xyz[3 * (size_t)i + 1];
		const double dz = out.aabb_center[2] - out.xyz[3 * (size_t)i + 2];
		const double r = dx * dx + dy * dy + dz * dz;
		if (r > r2) r2 = r;
	}
	out.aabb_radius = std::sqrt(r2);
	return true;
}

// [FCL-recalled] Quaternion3f::toRotation
void quat_to_rotation(const double q[4], double R[9]) {
	const double twoX = 2.0 * q[1], twoY = 2.0 * q[2], twoZ = 2.0 * q[3];
	const double twoWX = twoX * q[0], twoWY = twoY * q[0], twoWZ = twoZ * q[0];
	const double twoXX = twoX * q[1], twoXY = twoY * q[1], twoXZ = twoZ * q[1];
	const double twoYY = twoY * q[2], twoYZ = twoZ * q[2], twoZZ = twoZ * q[3];
	R[0] = 1.0 - (twoYY + twoZZ);
	R[1] = twoXY - twoWZ;
	R[2] = twoXZ + twoWY;
	R[3] = twoXY + twoWZ;
	R[4] = 1.0 - (twoXX + twoZZ);
	R[5] = twoYZ - twoWX;
	R[6] = twoXZ - twoWY;
	R[7] = twoYZ + twoWX;
	R[8] = 1.0 - (twoXX + twoYY);
}

}  // namespace mptcc
