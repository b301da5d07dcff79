// mptcc_api.cpp -- C ABI of libmptcc (include/mptcc.h): context, scene flattening, device replication,
// pose sharding over 1/2/4/8 GPUs, host<->device pipelining.  Host-side exact arithmetic (environment
// object transforms and world AABBs) restates fcl::Transform3f / fcl::CollisionObject [FCL-recalled
// C.1, C.5]; build with -ffp-contract=off.
#include <cuda_runtime.h>

#include <atomic>
#include <thread>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mptcc.h"
#include "bvh.hpp"
#include "device_layout.hpp"

using namespace mptcc;

namespace {

// ---- small host helpers -------------------------------------------------------------------------
inline float round_up_f32(double x) {
	float f = (float)x;
	if ((double)f < x) f = std::nextafterf(f, INFINITY);
	return f;
}

inline float round_down_f32(double x) {
	float f = (float)x;
	if ((double)f > x) f = std::nextafterf(f, -INFINITY);
	return f;
}

// IEEE binary16 helpers for the quantised nodes (device_layout.hpp, nodeq): exact half -> float, and float -> half
// rounded to nearest (ties away from zero are fine here: the error bound is what matters) or upward.
inline float half_to_float(uint16_t h) {
	const int e = (h >> 10) & 31, m = h & 1023;
	float v;
	if (e == 0) v = std::ldexp((float)m, -24);
	else if (e == 31) v = m ? NAN : INFINITY;
	else v = std::ldexp((float)(m + 1024), e - 25);
	return (h & 0x8000) ? -v : v;
}
inline uint16_t float_to_half(float f, bool up) {  // up: smallest half >= f (f >= 0); else nearest
	if (!(f == f)) return 0x7e00;
	const uint16_t sign = std::signbit(f) ? 0x8000 : 0;
	const float a = std::fabs(f);
	uint16_t h;  // truncation toward zero of |f|
	if (a >= 65504.0f) h = 0x7bff;
	else if (a < std::ldexp(1.0f, -24)) h = 0;
	else {
		int e;
		const float m = std::frexp(a, &e);  // a = m * 2^e, m in [0.5, 1)
		if (e - 1 < -14) h = (uint16_t)std::floor(std::ldexp(a, 24));
		else h = (uint16_t)(((e - 1 + 15) << 10) | (int)std::floor((m * 2.0f - 1.0f) * 1024.0f));
	}
	const float lo = half_to_float(h), hi = half_to_float((uint16_t)(h + 1));
	if (up) {
		if (lo < a) ++h;
	} else if (a - lo >= hi - a && hi != INFINITY) {
		++h;
	}
	return (uint16_t)(sign | h);
}

// [FCL-recalled] Quaternion3f::operator* / transform
inline void qmul(const double a[4], const double b[4], double o[4]) {
	const double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
	const double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
	const double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
	const double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
	o[0] = w;
	o[1] = x;
	o[2] = y;
	o[3] = z;
}

inline void quat_rotate(const double q[4], const double v[3], double o[3]) {
	const double p[4] = {0.0, v[0], v[1], v[2]};
	const double c[4] = {q[0], -q[1], -q[2], -q[3]};
	double t[4], r[4];
	qmul(q, p, t);
	qmul(t, c, r);
	o[0] = r[1];
	o[1] = r[2];
	o[2] = r[3];
}

// Quantised nodes (device_layout.hpp, nodeq): each component of a decoded axis is within kQAxisDelta of the true one
// (kQAxisCheck is what the host verifies in exact arithmetic; the rest is the kernel's fp32 evaluation).
constexpr double kQAxisDelta = 4e-6, kQAxisCheck = 2e-6;

// Host image of one ModelPack.
struct HostPack {
	std::vector<float4> node32;
	std::vector<uint4> nodeq;
	std::vector<float4> tri32;
	std::vector<double> node64;
	std::vector<double> tri64;
	std::vector<uint32_t> anc;
	std::vector<float> mu_chain;
	std::vector<DevModel> models;
	uint32_t max_depth = 0;
	bool q_ok = true;  // every node could be written as a quantised node (nodeq)

	void clear() { *this = HostPack(); }

	// Appends one model; returns its id.  bfs: number the device nodes breadth-first (root 0, the two children of a
	// node adjacent, parents before children level by level) instead of in FCL's allocation order, so that the top
	// levels of the tree are one contiguous block the narrowphase kernel can stage in shared memory with one bulk
	// copy.  Only the device ids change (links, ancestor chains); primitives and the fp64 node values do not.
	uint32_t add(const HostBvh &b, bool bfs = false) {
		const uint32_t nn0 = b.num_nodes();
		std::vector<uint32_t> order(nn0), newid(nn0);  // order[new] = FCL id, newid[FCL id] = new
		if (bfs) {
			uint32_t head = 0, tail = 0;
			order[tail++] = 0;
			while (head < tail) {
				const uint32_t v = order[head++];
				if (b.first_child[v] >= 0) {
					order[tail++] = (uint32_t)b.first_child[v];
					order[tail++] = (uint32_t)b.first_child[v] + 1u;
				}
			}
		} else {
			for (uint32_t i = 0; i < nn0; ++i) order[i] = i;
		}
		for (uint32_t i = 0; i < nn0; ++i) newid[order[i]] = i;
		DevModel m;
		std::memset(&m, 0, sizeof m);
		if ((node32.size() / 4) % 2 == 0) {  // node_base must be odd: pad one node slot
			for (int k = 0; k < 4; ++k) node32.push_back(make_float4(0, 0, 0, 0));
			for (int k = 0; k < 2; ++k) nodeq.push_back(make_uint4(0, 0, 0, 0));
			for (int k = 0; k < 16; ++k) node64.push_back(0.0);
		}
		m.node_base = (uint32_t)(node32.size() / 4);
		m.tri_base = (uint32_t)(tri32.size() / 3);
		m.num_tris = b.num_tris();
		m.depth = b.depth;
		max_depth = std::max(max_depth, b.depth);
		for (int k = 0; k < 3; ++k) {
			m.ctr[k] = b.aabb_center[k];
			m.aabb_min[k] = b.aabb_min[k];
			m.aabb_max[k] = b.aabb_max[k];
			m.aabb_center[k] = b.aabb_center[k];
		}
		m.aabb_radius = b.aabb_radius;
		double r2 = 0, rinf = 0;
		const uint32_t nv = (uint32_t)(b.xyz.size() / 3);
		for (uint32_t i = 0; i < nv; ++i) {
			double d2 = 0;
			for (int k = 0; k < 3; ++k) {
				const double d = b.xyz[3 * (size_t)i + k] - m.ctr[k];
				d2 += d * d;
				rinf = std::max(rinf, std::fabs(d));
			}
			r2 = std::max(r2, d2);
		}
		m.r2 = round_up_f32(std::sqrt(r2) * (1.0 + 1e-6));
		m.rinf = round_up_f32(rinf * (1.0 + 1e-6));

		const uint32_t nn = b.num_nodes();
		std::vector<double> node_mu(nn);
		for (uint32_t i = 0; i < nn; ++i) {
			const double *C = &b.center[3 * (size_t)i];
			const double *E = &b.extent[3 * (size_t)i];
			const double emin = std::min(E[0], std::min(E[1], E[2]));
			const double emax = std::max(E[0], std::max(E[1], E[2]));
			const double emid = std::max(std::min(E[0], E[1]), std::min(std::max(E[0], E[1]), E[2]));
			(void)emax;
			node_mu[i] = emin + emid;
			m.cmax1 = std::max(m.cmax1, std::fabs(C[0]) + std::fabs(C[1]) + std::fabs(C[2]));
			m.esum = std::max(m.esum, E[0] + E[1] + E[2]);
		}
		for (uint32_t ni = 0; ni < nn; ++ni) {
			const uint32_t i = order[ni];
			const double *A = &b.axis[9 * (size_t)i];
			const double *C = &b.center[3 * (size_t)i];
			const double *E = &b.extent[3 * (size_t)i];
			for (int k = 0; k < 3; ++k)
				node32.push_back(make_float4((float)A[3 * k], (float)A[3 * k + 1], (float)A[3 * k + 2], (float)(C[k] - m.ctr[k])));
			float4 last;
			last.x = round_up_f32(E[0] * (1.0 + 1e-6));
			last.y = round_up_f32(E[1] * (1.0 + 1e-6));
			last.z = round_up_f32(E[2] * (1.0 + 1e-6));
			const int32_t link = b.first_child[i] >= 0 ? (int32_t)newid[(uint32_t)b.first_child[i]] : b.first_child[i];
			std::memcpy(&last.w, &link, 4);
			node32.push_back(last);
			{
				// quantised node (32 bytes, layout: device_layout.hpp): centre fp32, extents binary16 (rounded up), the
				// orientation as a unit quaternion with four 22-bit fixed-point components, 24-bit link.
				// The rows of the rotation matrix of the quaternion are the box axes; an OBB does not change when an axis
				// is negated, so a left-handed axis triple is stored with axis 2 flipped.
				double M[9];
				for (int k = 0; k < 9; ++k) M[k] = A[k];
				const double det = M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
				if (det < 0.0)
					for (int k = 6; k < 9; ++k) M[k] = -M[k];
				double q[4];  // w x y z
				{
					const double tr = M[0] + M[4] + M[8];
					if (tr > 0.0) {
						const double r = std::sqrt(tr + 1.0), f = 0.5 / r;
						q[0] = 0.5 * r;
						q[1] = (M[7] - M[5]) * f;
						q[2] = (M[2] - M[6]) * f;
						q[3] = (M[3] - M[1]) * f;
					} else {
						int i = 0;
						if (M[4] > M[0]) i = 1;
						if (M[8] > M[4 * i]) i = 2;
						const int j = (i + 1) % 3, k2 = (i + 2) % 3;
						const double r = std::sqrt(M[4 * i] - M[4 * j] - M[4 * k2] + 1.0), f = 0.5 / r;
						q[1 + i] = 0.5 * r;
						q[0] = (M[3 * k2 + j] - M[3 * j + k2]) * f;
						q[1 + j] = (M[3 * i + j] + M[3 * j + i]) * f;
						q[1 + k2] = (M[3 * i + k2] + M[3 * k2 + i]) * f;
					}
					const double nq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
					for (int k = 0; k < 4; ++k) q[k] = nq > 0.0 && std::isfinite(nq) ? q[k] / nq : (k == 0 ? 1.0 : 0.0);
				}
				uint32_t Q[4];
				double qd[4];  // the components the kernel decodes: Q * 2^-21 - 1
				for (int k = 0; k < 4; ++k) {
					const double v = std::floor((q[k] + 1.0) * 2097152.0 + 0.5);
					Q[k] = (uint32_t)std::min(4194303.0, std::max(0.0, v));
					qd[k] = (double)Q[k] / 2097152.0 - 1.0;
				}
				{
					// the axes as the kernel reconstructs them (exact arithmetic here; its fp32 evaluation is within 1e-6 of
					// this) must be within kQAxisCheck of the true ones, or the pack cannot use quantised nodes
					const double w = qd[0], x = qd[1], y = qd[2], z = qd[3];
					const double Rq[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
					                      2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
					                      2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
					for (int k = 0; k < 9; ++k)
						if (!(std::fabs(Rq[k] - M[k]) <= kQAxisCheck)) q_ok = false;
				}
				const double esum = E[0] + E[1] + E[2];
				uint16_t he[3];
				for (int k = 0; k < 3; ++k) {
					// widened so that the box spanned by the decoded axes contains FCL's box
					const double e = (E[k] + 1.7320508075688772 * kQAxisDelta * esum) * (1.0 + 1e-6);
					if (!(e <= 65000.0)) q_ok = false;
					he[k] = float_to_half(round_up_f32(e), true);
				}
				if (link >= (1 << 23) || link < -(1 << 23)) q_ok = false;
				const uint32_t raw = (uint32_t)link & 0xffffffu;  // 24-bit two's complement
				const float cq[3] = {(float)(C[0] - m.ctr[0]), (float)(C[1] - m.ctr[1]), (float)(C[2] - m.ctr[2])};
				uint4 q0, q1;
				std::memcpy(&q0.x, &cq[0], 4);
				std::memcpy(&q0.y, &cq[1], 4);
				std::memcpy(&q0.z, &cq[2], 4);
				q0.w = he[0] | ((uint32_t)he[1] << 16);
				q1.x = he[2] | ((raw & 0xffffu) << 16);
				q1.y = Q[0] | (Q[1] << 22);
				q1.z = (Q[1] >> 10) | (Q[2] << 12);
				q1.w = (Q[2] >> 20) | (Q[3] << 2) | (((raw >> 16) & 0xffu) << 24);
				nodeq.push_back(q0);
				nodeq.push_back(q1);
			}

			for (int k = 0; k < 9; ++k) node64.push_back(A[k]);
			for (int k = 0; k < 3; ++k) node64.push_back(C[k]);
			for (int k = 0; k < 3; ++k) node64.push_back(E[k]);
			node64.push_back(E[0] * E[0] + E[1] * E[1] + E[2] * E[2]);  // Vec3f::sqrLength, OBB::size()

		}
		const uint32_t nt = b.num_tris();
		// ancestor chains: anc[anc_base + prim * depth + level] = node id at that level of the path to prim
		m.anc_base = anc.size();
		{
			std::vector<int32_t> parent(nn, -1);
			for (uint32_t i = 0; i < nn; ++i)
				if (b.first_child[i] >= 0) {
					parent[b.first_child[i]] = (int32_t)i;
					parent[b.first_child[i] + 1] = (int32_t)i;
				}
			anc.resize(anc.size() + (size_t)nt * b.depth, 0xffffffffu);
			const size_t mu_base = mu_chain.size();
			mu_chain.resize(mu_base + nt, 0.f);
			std::vector<uint32_t> chain;
			for (uint32_t i = 0; i < nn; ++i) {
				if (b.first_child[i] >= 0) continue;
				const uint32_t prim = (uint32_t)(-(b.first_child[i] + 1));
				chain.clear();
				for (int32_t v = (int32_t)i; v >= 0; v = parent[v]) chain.push_back((uint32_t)v);
				uint32_t *dst = &anc[m.anc_base + (size_t)prim * b.depth];
				for (size_t l = 0; l < chain.size(); ++l) dst[l] = newid[chain[chain.size() - 1 - l]];
				double mu = DBL_MAX;
				for (uint32_t v : chain) mu = std::min(mu, node_mu[v]);
				mu_chain[mu_base + prim] = round_down_f32(mu);
			}
		}
		for (uint32_t t = 0; t < nt; ++t) {
			for (int c = 0; c < 3; ++c) {
				const double *v = &b.xyz[3 * (size_t)b.tri[3 * (size_t)t + c]];
				tri32.push_back(make_float4((float)(v[0] - m.ctr[0]), (float)(v[1] - m.ctr[1]), (float)(v[2] - m.ctr[2]), 0.f));
				tri64.push_back(v[0]);
				tri64.push_back(v[1]);
				tri64.push_back(v[2]);
			}
		}
		models.push_back(m);
		return (uint32_t)(models.size() - 1);
	}
};

// Broadphase: Morton order of the objects + threaded binary tree over clusters of kClusterSize objects
// (layout: device_layout.hpp).  Replaces FCL's DynamicAABBTreeCollisionManager for candidate
// generation; the candidates it yields are re-tested with FCL's exact fp64 AABB::overlap, so the set
// of (object, link) pairs that reach the narrowphase is FCL's [FCL-recalled C.6].
uint32_t morton10(double x) {  // x in [0,1)
	uint32_t v = (uint32_t)std::min(1023.0, std::max(0.0, x * 1024.0));
	v = (v | (v << 16)) & 0x030000FFu;
	v = (v | (v << 8)) & 0x0300F00Fu;
	v = (v | (v << 4)) & 0x030C30C3u;
	v = (v | (v << 2)) & 0x09249249u;
	return v;
}

struct BpBuilder {
	const std::vector<DevObject> &objs;
	std::vector<uint4> nodes;
	float origin[3], inv_cell[3];  // grid of the 16-bit node boxes: q(x) = (x - origin) * inv_cell
	explicit BpBuilder(const std::vector<DevObject> &o) : objs(o) {
		double lo[3], hi[3];
		cluster_box(0, (objs.size() + kClusterSize - 1) / kClusterSize, lo, hi);
		for (int k = 0; k < 3; ++k) {
			origin[k] = round_down_f32(lo[k]);
			const double extent = hi[k] - (double)origin[k];
			// the scene spans 65 000 of the 65 536 cells: rounding never pushes a box past the last cell
			inv_cell[k] = extent > 0.0 && std::isfinite(extent) ? (float)(65000.0 / extent) : 1.0f;
		}
	}

	void cluster_box(size_t c0, size_t c1, double lo[3], double hi[3]) const {
		for (int k = 0; k < 3; ++k) {
			lo[k] = DBL_MAX;
			hi[k] = -DBL_MAX;
		}
		const size_t b = c0 * kClusterSize, e = std::min(objs.size(), c1 * kClusterSize);
		for (size_t i = b; i < e; ++i)
			for (int k = 0; k < 3; ++k) {
				lo[k] = std::min(lo[k], objs[i].aabb[k]);
				hi[k] = std::max(hi[k], objs[i].aabb[3 + k]);
			}
	}

	// cell index of x, rounded outward (down for a box minimum, up for a maximum); fp64 here, the 1e-6
	// covers its rounding.  Non-finite boxes cover the whole grid.
	uint32_t cell(double x, int k, bool up) const {
		const double q = (x - (double)origin[k]) * (double)inv_cell[k];
		if (!(q == q)) return up ? 65535u : 0u;
		const double r = up ? std::ceil(q + 1e-6) : std::floor(q - 1e-6);
		return (uint32_t)std::min(65535.0, std::max(0.0, r));
	}

	// pre-order emission of the subtree over clusters [c0, c1)
	void emit(size_t c0, size_t c1) {
		const size_t me = nodes.size();
		double lo[3], hi[3];
		cluster_box(c0, c1, lo, hi);
		uint4 nd;
		nd.x = cell(lo[0], 0, false) | (cell(lo[1], 1, false) << 16);
		nd.y = cell(lo[2], 2, false) | (cell(hi[0], 0, true) << 16);
		nd.z = cell(hi[1], 1, true) | (cell(hi[2], 2, true) << 16);
		nd.w = 0u;
		nodes.push_back(nd);
		if (c1 - c0 == 1) {
			nodes[me].w = 0x80000000u | (uint32_t)c0;  // leaf: its escape is simply the next node
		} else {
			const size_t mid = c0 + (c1 - c0) / 2;
			emit(c0, mid);
			emit(mid, c1);
			nodes[me].w = (uint32_t)nodes.size();  // escape
		}
	}
};

void sort_objects_morton(std::vector<DevObject> &objs) {
	if (objs.size() < 2) return;
	double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
	for (const auto &o : objs)
		for (int k = 0; k < 3; ++k) {
			const double c = 0.5 * (o.aabb[k] + o.aabb[3 + k]);
			lo[k] = std::min(lo[k], c);
			hi[k] = std::max(hi[k], c);
		}
	std::vector<std::pair<uint32_t, uint32_t>> key(objs.size());
	for (size_t i = 0; i < objs.size(); ++i) {
		uint32_t code = 0;
		for (int k = 0; k < 3; ++k) {
			const double c = 0.5 * (objs[i].aabb[k] + objs[i].aabb[3 + k]);
			const double span = hi[k] - lo[k];
			const double u = span > 0 ? (c - lo[k]) / span : 0.0;
			code |= morton10(u * (1023.0 / 1024.0)) << k;
		}
		key[i] = std::make_pair(code, (uint32_t)i);
	}
	std::stable_sort(key.begin(), key.end());
	std::vector<DevObject> sorted(objs.size());
	for (size_t i = 0; i < objs.size(); ++i) sorted[i] = objs[key[i].second];
	objs.swap(sorted);
}

template <class T>
struct DevBuf {
	T *p = nullptr;
	size_t n = 0;
	cudaError_t upload(const std::vector<T> &h, cudaStream_t s) {
		cudaError_t e = reserve(h.size());
		if (e != cudaSuccess) return e;
		if (h.empty()) return cudaSuccess;
		return cudaMemcpyAsync(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, s);
	}
	cudaError_t reserve(size_t count) {
		if (count <= n && p) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr;
		n = 0;
		const size_t bytes = sizeof(T) * std::max<size_t>(count, 1);
		cudaError_t e = cudaMalloc((void **)&p, bytes);
		if (e == cudaSuccess) n = std::max<size_t>(count, 1);
		return e;
	}
	void release() {
		if (p) cudaFree(p);
		p = nullptr;
		n = 0;
	}
};

struct DevPack {
	DevBuf<float4> node32, tri32;
	DevBuf<uint4> nodeq;
	DevBuf<double> node64, tri64;
	DevBuf<uint32_t> anc;
	DevBuf<float> mu_chain;
	DevBuf<DevModel> models;
	cudaError_t upload(const HostPack &h, cudaStream_t s) {
		cudaError_t e;
		if ((e = node32.upload(h.node32, s)) != cudaSuccess) return e;
		if ((e = nodeq.upload(h.nodeq, s)) != cudaSuccess) return e;
		if ((e = tri32.upload(h.tri32, s)) != cudaSuccess) return e;
		if ((e = node64.upload(h.node64, s)) != cudaSuccess) return e;
		if ((e = tri64.upload(h.tri64, s)) != cudaSuccess) return e;
		if ((e = anc.upload(h.anc, s)) != cudaSuccess) return e;
		if ((e = mu_chain.upload(h.mu_chain, s)) != cudaSuccess) return e;
		return models.upload(h.models, s);
	}
	ModelPack view() const { return ModelPack{node32.p, tri32.p, node64.p, tri64.p, anc.p, mu_chain.p, models.p, nodeq.p}; }
	void release() {
		node32.release();
		nodeq.release();
		tri32.release();
		node64.release();
		tri64.release();
		anc.release();
		mu_chain.release();
		models.release();
	}
};

#ifndef MPTCC_PIPE
#define MPTCC_PIPE 4
#endif
constexpr int kPipe = MPTCC_PIPE;  // staging sets (pose buffer, verdict words, candidate list, stream) per device: the sub-batches
                                   // of a call rotate over them, so the copies of the next sub-batches are in flight
                                   // while the kernels of the previous ones run
constexpr size_t kSmallBytes = 256 << 10;  // pose bytes up to which a blocking call takes the zero-copy path
constexpr int kInFlight = 4;  // mptcc_submit_poses calls that may be outstanding before submit waits for the oldest

// Buffers of the split pipeline (split_kernels.cu) for one launch slot; grow-only.
struct SplitBufs {
	DevBuf<float4> rec;
	DevBuf<double> rel;
	DevBuf<uint32_t> obj, pose;
	DevBuf<uint2> base;        // node_base / tri_base of each candidate's environment model
	DevBuf<float> cert;        // path-certificate threshold of each candidate
	DevBuf<uint4> xq;          // exact-pass queue (np2_kernels.cu)
	DevBuf<unsigned int> ctl;  // [0] cand_total, [1] overflow flags (1 = list full, 2 = non-rigid pose), [2] narrowphase
	                           // batch counter, [3] exact-queue length
	void release() {
		rec.release();
		rel.release();
		obj.release();
		pose.release();
		base.release();
		cert.release();
		xq.release();
		ctl.release();
	}
};

struct Device {
	int id = 0;
	int sm_count = 0;
	size_t l2_persist_max = 0, l2_window_max = 0;  // persisting-L2 carve-out / access policy window limits of the device
	cudaStream_t stream[kPipe] = {};
	cudaEvent_t ev0[kPipe] = {}, ev1[kPipe] = {};
	DevPack env, agent;
	DevBuf<DevObject> objects;
	DevBuf<double> obj_aabb;
	DevBuf<uint4> bp_nodes;
	DevBuf<uint32_t> link_model;
	DevBuf<DevObject> self_objects;  // one per link (model id only): the "objects" of the self-collision kernel
	DevBuf<uint2> grid_cells;        // broadphase grid (device_layout.hpp)
	DevBuf<GridEntry> grid_entries;
	DevBuf<float4> agent_swz;        // link model 0's traversal nodes, bank-swizzled for the np3 kernel's shared-memory copy
	DevBuf<double> poses[kPipe];
	DevBuf<uint32_t> bits[kPipe];
	DevBuf<uint32_t> work[kPipe];
	DevBuf<unsigned long long> counters;
	DevBuf<uint32_t> work_user;  // work counter of the device-buffer entry point
	SplitBufs split[kPipe + 2];  // staging sets, device-buffer entry point, edge entry point
	cudaEvent_t phase_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // option time_phases: around the pipeline kernels
	// on-device edge interpolation (mptcc_check_omni_edges)
	DevBuf<double> edge_in;
	DevBuf<unsigned long long> edge_off;
	DevBuf<double> edge_poses;
	DevBuf<uint32_t> edge_bits;
	DevBuf<unsigned char> edge_safe;
	DevBuf<uint32_t> edge_work;
	float last_ms = 0.f;
	// planner-sized calls: poses / verdict words in pinned host memory the kernel reads and writes in place
	double *small_in = nullptr, *small_in_dev = nullptr;     // host address / the device's address of the same memory
	uint32_t *small_out = nullptr, *small_out_dev = nullptr;
	// per outstanding submit: completion events of the two streams, overflow flags read back (pinned)
	cudaEvent_t done[kInFlight][kPipe] = {};
	unsigned int *ctl_h[kInFlight] = {};
	size_t used[kInFlight] = {};
	bool split_ran[kPipe] = {};  // the last job enqueued on stream k went through the split pipeline
};

}  // namespace

struct mptcc_ctx {
	std::vector<Device> devs;
	HostPack env_h, agent_h;
	std::vector<DevObject> objects_h;
	std::vector<double> obj_aabb_h;
	std::vector<uint4> bp_nodes_h;
	float bp_origin[3] = {0.f, 0.f, 0.f}, bp_inv_cell[3] = {1.f, 1.f, 1.f};
	std::vector<uint32_t> link_model_h;
	std::vector<DevObject> self_objects_h;
	std::vector<float4> agent_swz_h;
	std::vector<uint2> grid_cells_h;  // broadphase grid: empty = not available (the tree is walked instead)
	std::vector<GridEntry> grid_entries_h;
	float grid_origin[3] = {0.f, 0.f, 0.f}, grid_inv[3] = {1.f, 1.f, 1.f};
	int grid_n[3] = {1, 1, 1};
	bool env_set = false, agent_set = false;
	bool env_q_ok = true, agent_q_ok = true;  // every model small enough for the 24-bit links of the quantised nodes
	std::string err;
	bool count = false;
	long opt_ppw = 0;           // 0 = auto
	long opt_ctas_per_sm = 0;   // 0 = occupancy maximum
	long opt_sub_batch = 1 << 19;    // pose-sets per piece of a host-buffer call (a call of >= 2^16 pose-sets is cut into at least two)
	long opt_always_replay = 0;
	long opt_time_phases = 0;        // record CUDA events around the kernels of the split pipeline
	long opt_split = 1;              // two-kernel pipeline for large single-link batches
	long opt_split_min = 1 << 15;    // ... from this many pose-sets per launch
	long opt_split_cap = 3;          // candidate list capacity, per pose-set (grows after an overflow, see check_poses)
	long opt_no_steal = 0;
	long opt_steal_bottom = 0;
	long opt_near_first = 1;
	long opt_long_batches = -1;      // narrowphase batches of 128 instead of 96 candidates: -1 = by launch size
	long opt_self_collision = 0;     // checkSelfCollision = true of MeshHandler::isInCollision
	long opt_np2 = 0;                // 1 = second-generation narrowphase (np2_kernels.cu): same speed on C5, see DESIGN.md section 5
	long opt_expand4 = 0;            // np2: 1 = expand both trees per step (four tests); measured 37 % slower on C5
	long opt_l2_persist = -1;        // persisting-L2 access window over the environment's traversal nodes during the narrowphase: 1 / 0, -1 = when it pays (enqueue_kernel)
	long opt_qnode = 1;              // np2/np3 traverse the 32-byte quantised nodes (one 256-bit load per node); 0 = the 64-byte fp32 nodes
	long opt_agent_top = -1;         // agent nodes staged in shared memory by the np2 kernel: -1 = as many as fit (511)
	long opt_first_half = 1;         // a host-buffer call that finds the device idle cuts its first piece in two
	long opt_grid = 1;               // grid broadphase for single-link agents (0 = always walk the cluster tree)
	long opt_np3 = 1;                // third-generation narrowphase (np3_kernels.cu, streaming lanes): the default; 0 = the round-1 kernel
	                                 // (split_kernels.cu), or np2_kernels.cu with np2 = 1
	mptcc_counters ctr{};
	uint64_t submitted = 0, completed = 0;  // tickets of mptcc_submit_poses
	int ctas_per_sm_cached = -1;
	bool ctas_cached_count = false;
};

struct mptcc_bvh {
	HostBvh b;
};

namespace {

int fail(mptcc_ctx *c, int code, const std::string &msg) {
	if (c) c->err = msg;
	return code;
}

int cuda_fail(mptcc_ctx *c, cudaError_t e, const char *what) {
	cudaGetLastError();
	return fail(c, MPTCC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CU(call)                                                 \
	do {                                                         \
		cudaError_t e__ = (call);                                \
		if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
	} while (0)

unsigned stack_guard_for(const mptcc_ctx *c) { return c->env_h.max_depth + c->agent_h.max_depth + 2; }

int grid_for(mptcc_ctx *ctx, const Device &d) {
	if (ctx->ctas_per_sm_cached < 0 || ctx->ctas_cached_count != ctx->count) {
		ctx->ctas_per_sm_cached = collide_max_ctas_per_sm(ctx->count);
		ctx->ctas_cached_count = ctx->count;
	}
	int per_sm = ctx->ctas_per_sm_cached;
	if (per_sm < 1) per_sm = 1;
	if (ctx->opt_ctas_per_sm > 0 && ctx->opt_ctas_per_sm < per_sm) per_sm = (int)ctx->opt_ctas_per_sm;
	return d.sm_count * per_sm;
}

unsigned pick_ppw_log2(const mptcc_ctx *ctx, uint64_t n, int grid) {
	if (ctx->opt_ppw > 0) {
		unsigned l = 0;
		while ((1l << (l + 1)) <= ctx->opt_ppw && l < 6) ++l;
		return l;
	}
	// enough chunks for ~2 per resident warp, else fewer poses per warp
	const uint64_t warps = (uint64_t)grid * kWarpsPerCta;
	unsigned l = 5;  // 64 pose slots per warp are supported (option poses_per_warp) but measured slower
	while (l > 0 && (n >> l) < 2 * warps) --l;
	return l;
}

int upload_scene_part(mptcc_ctx *ctx, bool env) {
	for (auto &d : ctx->devs) {
		CU(cudaSetDevice(d.id));
		if (env) {
			CU(d.env.upload(ctx->env_h, d.stream[0]));
			CU(d.objects.upload(ctx->objects_h, d.stream[0]));
			CU(d.obj_aabb.upload(ctx->obj_aabb_h, d.stream[0]));
			CU(d.bp_nodes.upload(ctx->bp_nodes_h, d.stream[0]));
		} else {
			CU(d.agent.upload(ctx->agent_h, d.stream[0]));
			CU(d.link_model.upload(ctx->link_model_h, d.stream[0]));
			CU(d.self_objects.upload(ctx->self_objects_h, d.stream[0]));
			CU(d.agent_swz.upload(ctx->agent_swz_h, d.stream[0]));
		}
		CU(cudaStreamSynchronize(d.stream[0]));
	}
	return MPTCC_OK;
}

// Broadphase grid for a single-link agent (device_layout.hpp).  Needs the environment (object boxes) and the agent
// (largest pose box); rebuilt whenever either changes.  Leaves grid_cells_h empty when the grid does not apply.
int build_grid(mptcc_ctx *ctx) {
	ctx->grid_cells_h.clear();
	ctx->grid_entries_h.clear();
	const auto upload = [&]() -> int {
		for (auto &d : ctx->devs) {
			CU(cudaSetDevice(d.id));
			CU(d.grid_cells.upload(ctx->grid_cells_h, d.stream[0]));
			CU(d.grid_entries.upload(ctx->grid_entries_h, d.stream[0]));
			CU(cudaStreamSynchronize(d.stream[0]));
		}
		return MPTCC_OK;
	};
	if (ctx->objects_h.empty() || ctx->link_model_h.size() != 1 || ctx->agent_h.models.empty()) return upload();
	const DevModel &am = ctx->agent_h.models[ctx->link_model_h[0]];
	double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX}, amax = 0.0;
	for (const auto &o : ctx->objects_h)
		for (int k = 0; k < 3; ++k) {
			if (!std::isfinite(o.aabb[k]) || !std::isfinite(o.aabb[3 + k])) return upload();
			lo[k] = std::min(lo[k], o.aabb[k]);
			hi[k] = std::max(hi[k], o.aabb[3 + k]);
			amax = std::max(amax, std::max(std::fabs(o.aabb[k]), std::fabs(o.aabb[3 + k])));
		}
	// Every pose box is at most this wide: the sphere box of a rotated pose is 2 radius wide and contains the tight
	// box of an unrotated one.  The slack covers the fp32 rounding of the coordinates inside grid_cell().
	if (!std::isfinite(am.aabb_radius)) return upload();
	const double side = 2.0 * am.aabb_radius * (1.0 + 1e-4) + 8.0 * (amax + 2.0 * am.aabb_radius) * 1.2e-7 + 1e-30;
	for (int nmax = 128; nmax >= 1; nmax /= 2) {
		for (int k = 0; k < 3; ++k) {
			const double ext = hi[k] - lo[k];
			const double cell = std::max(side, ext / nmax);
			ctx->grid_origin[k] = round_down_f32(lo[k]);
			ctx->grid_inv[k] = round_down_f32(1.0 / cell);
			ctx->grid_n[k] = std::max(1, std::min(nmax, (int)std::ceil(ext / cell) + 1));
			if (!(ctx->grid_inv[k] > 0.f) || !std::isfinite(ctx->grid_inv[k])) return upload();
		}
		const int n0 = ctx->grid_n[0], n1 = ctx->grid_n[1], n2 = ctx->grid_n[2];
		const size_t ncell = (size_t)n0 * n1 * n2;
		std::vector<uint32_t> cnt(ncell + 1, 0u);
		// blocks (i, j, k) whose cells [i, i+1] touch the object's cells [a, b]: i in [a - 1, b]
		int rng[6];
		const auto range_of = [&](const DevObject &o) {
			for (int k = 0; k < 3; ++k) {
				const int a = grid_cell(o.aabb[k], ctx->grid_origin[k], ctx->grid_inv[k], ctx->grid_n[k]);
				const int b = grid_cell(o.aabb[3 + k], ctx->grid_origin[k], ctx->grid_inv[k], ctx->grid_n[k]);
				rng[k] = std::max(0, a - 1);
				rng[3 + k] = b;
			}
		};
		size_t total = 0;
		for (const auto &o : ctx->objects_h) {
			range_of(o);
			total += (size_t)(rng[3] - rng[0] + 1) * (rng[4] - rng[1] + 1) * (rng[5] - rng[2] + 1);
		}
		if (total > (size_t)(3u << 19) && nmax > 1) continue;  // more than 96 MB of entries: a coarser grid
		for (const auto &o : ctx->objects_h) {
			range_of(o);
			for (int i = rng[0]; i <= rng[3]; ++i)
				for (int j = rng[1]; j <= rng[4]; ++j)
					for (int k = rng[2]; k <= rng[5]; ++k) ++cnt[((size_t)i * n1 + j) * n2 + k + 1];
		}
		for (size_t c = 0; c < ncell; ++c) cnt[c + 1] += cnt[c];
		ctx->grid_cells_h.resize(ncell);
		for (size_t c = 0; c < ncell; ++c) ctx->grid_cells_h[c] = make_uint2(cnt[c], cnt[c + 1] - cnt[c]);
		ctx->grid_entries_h.assign(total, GridEntry{});
		std::vector<uint32_t> fill(cnt.begin(), cnt.end() - 1);
		for (size_t oi = 0; oi < ctx->objects_h.size(); ++oi) {
			const DevObject &o = ctx->objects_h[oi];
			range_of(o);
			for (int i = rng[0]; i <= rng[3]; ++i)
				for (int j = rng[1]; j <= rng[4]; ++j)
					for (int k = rng[2]; k <= rng[5]; ++k) {
						GridEntry &g = ctx->grid_entries_h[fill[((size_t)i * n1 + j) * n2 + k]++];
						for (int a = 0; a < 6; ++a) g.aabb[a] = o.aabb[a];
						g.obj = oi;
						g.pad = 0;
					}
		}
		break;
	}
	return upload();
}

// Outstanding submits and device-buffer launches still read the scene arrays: a set_* call (which may overwrite
// them in place) or an option that changes what later launches do first waits for all of it.
int drain(mptcc_ctx *ctx) {
	if (ctx->submitted != ctx->completed) {
		const int rc = mptcc_wait(ctx, 0);
		if (rc != MPTCC_OK) return rc;
	}
	for (auto &d : ctx->devs) {
		CU(cudaSetDevice(d.id));
		CU(cudaDeviceSynchronize());
	}
	return MPTCC_OK;
}

int check_ready(mptcc_ctx *ctx) {
	if (!ctx) return MPTCC_ERR_INVALID;
	if (!ctx->env_set) return fail(ctx, MPTCC_ERR_STATE, "environment not set (mptcc_set_environment)");
	if (!ctx->agent_set) return fail(ctx, MPTCC_ERR_STATE, "agent not set (mptcc_set_agent)");
	if (ctx->devs.empty()) return fail(ctx, MPTCC_ERR_NO_DEVICE, "context has no CUDA device");
	if (ctx->env_h.max_depth > 64 || ctx->agent_h.max_depth > 64)
		return fail(ctx, MPTCC_ERR_LIMIT, "BVH deeper than 64 levels is not supported by the exact path replay");
	if (stack_guard_for(ctx) > (unsigned)kLaneStackCap)
		return fail(ctx, MPTCC_ERR_LIMIT, "BVH depths exceed the per-lane traversal stack (depth1 + depth2 + 2 <= 64)");
	if (ctx->opt_self_collision && 2u * ctx->agent_h.max_depth + 2u > (unsigned)kLaneStackCap)
		return fail(ctx, MPTCC_ERR_LIMIT, "agent BVH too deep for the self-collision traversal (2 * depth + 2 <= 64)");
	return MPTCC_OK;
}

void fill_params(mptcc_ctx *ctx, Device &d, KernelParams &kp) {
	kp.env = d.env.view();
	kp.agent = d.agent.view();
	kp.objects = d.objects.p;
	kp.obj_aabb = d.obj_aabb.p;
	kp.n_objects_pad = (uint32_t)(ctx->obj_aabb_h.size() / 6);
	kp.bp_nodes = d.bp_nodes.p;
	kp.n_bp_nodes = (uint32_t)ctx->bp_nodes_h.size();
	for (int k = 0; k < 3; ++k) {
		kp.bp_origin[k] = ctx->bp_origin[k];
		kp.bp_inv_cell[k] = ctx->bp_inv_cell[k];
	}
	kp.link_model = d.link_model.p;
	kp.n_objects = (uint32_t)ctx->objects_h.size();
	kp.n_links = (uint32_t)ctx->link_model_h.size();
	kp.counters = ctx->count ? d.counters.p : nullptr;
	kp.always_replay = ctx->opt_always_replay ? 1u : 0u;
	kp.grid_cells = ctx->opt_grid && !ctx->grid_cells_h.empty() ? d.grid_cells.p : nullptr;
	kp.grid_entries = d.grid_entries.p;
	for (int k = 0; k < 3; ++k) {
		kp.grid_origin[k] = ctx->grid_origin[k];
		kp.grid_inv[k] = ctx->grid_inv[k];
		kp.grid_n[k] = ctx->grid_n[k];
	}
	kp.cand_rec = nullptr;
	kp.cand_rel = nullptr;
	kp.cand_obj = nullptr;
	kp.cand_pose = nullptr;
	kp.cand_total = nullptr;
	kp.overflow = nullptr;
	kp.cand_cap = 0;
	kp.run_if = nullptr;
	kp.work_counter2 = nullptr;
	kp.no_steal = ctx->opt_no_steal ? 1u : 0u;
	kp.steal_bottom = ctx->opt_steal_bottom ? 1u : 0u;
	kp.near_first = ctx->opt_near_first ? 1u : 0u;
}

// Candidate list buffers of the split pipeline for launches of up to n pose-sets.
int reserve_split(mptcc_ctx *ctx, SplitBufs &sb, uint64_t n) {
	if (!ctx->opt_split || ctx->link_model_h.size() != 1 || n < (uint64_t)ctx->opt_split_min) return MPTCC_OK;
	const uint64_t cap64 = std::min<uint64_t>(n * (uint64_t)ctx->opt_split_cap, 0xfffffff0ull);
	CU(sb.rec.reserve(4 * cap64));
	CU(sb.rel.reserve(12 * cap64));
	CU(sb.obj.reserve(cap64));
	CU(sb.pose.reserve(cap64));
	CU(sb.base.reserve(cap64));
	CU(sb.cert.reserve(cap64));
	CU(sb.xq.reserve(cap64));
	CU(sb.ctl.reserve(4));
	return MPTCC_OK;
}

// Entry points that do not go through mptcc_wait (device buffers, on-device edges) look at the overflow flag of their
// candidate list after a synchronise, so that a dense scene gets a longer list from the next launch on there too.
// The stream the list was used on must be idle.
void note_overflow(mptcc_ctx *ctx, Device &d, SplitBufs &sb) {
	(void)d;
	if (!sb.ctl.p) return;
	unsigned int h[4] = {0, 0, 0, 0};
	if (cudaMemcpy(h, sb.ctl.p, sizeof h, cudaMemcpyDeviceToHost) != cudaSuccess) {
		cudaGetLastError();
		return;
	}
	if ((h[1] & 1u) && ctx->opt_split_cap < 64) ctx->opt_split_cap = std::min<long>(64, ctx->opt_split_cap * 2);
	if (h[1]) cudaMemset(sb.ctl.p + 1, 0, sizeof(unsigned int));  // seen: a stale flag must not grow the list again
}

// Enqueue (memsets +) kernels for n pose-sets already on the device.  Large single-link batches take the
// split pipeline (broadphase kernel -> candidate list -> narrowphase kernel, then the fused kernel as a
// conditional overflow fall-back); everything else the fused kernel.
// checkSelfCollision: after the environment kernels of a launch, OR the link-vs-link verdicts into its words.
int enqueue_self(mptcc_ctx *ctx, Device &d, const KernelParams &kp_env, uint32_t *d_work, int grid, cudaStream_t s) {
	if (!ctx->opt_self_collision || ctx->link_model_h.size() < 2) return MPTCC_OK;
	KernelParams kp = kp_env;
	kp.env = kp.agent;
	kp.objects = d.self_objects.p;
	kp.n_objects = (uint32_t)ctx->link_model_h.size();
	kp.run_if = nullptr;
	kp.ppw_log2 = 5;
	kp.work_counter = d_work;  // (the environment launch may have run without a counter)
	CU(cudaMemsetAsync(kp.work_counter, 0, sizeof(uint32_t), s));
	const uint64_t need = ((kp.n_pose_sets + 31) / 32 + kWarpsPerCta - 1) / kWarpsPerCta;
	const int g = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)grid, need));
	const int e = launch_self_collide(kp, g, ctx->count, s);
	if (e != 0) return cuda_fail(ctx, (cudaError_t)e, "self_collide_kernel launch");
	ctx->ctr.n_kernel_launches++;
	return MPTCC_OK;
}

int enqueue_kernel(mptcc_ctx *ctx, Device &d, uint64_t n, const double *d_poses, int fmt, uint32_t *d_bits,
                   uint32_t *d_work, cudaStream_t s, SplitBufs *sb = nullptr, bool whole_words = false,
                   bool *ran_split = nullptr) {
	if (ran_split) *ran_split = false;
	KernelParams kp;
	fill_params(ctx, d, kp);
	const int grid = grid_for(ctx, d);
	kp.poses = d_poses;
	kp.pose_format = (uint32_t)fmt;
	kp.n_pose_sets = n;
	kp.verdict_bits = d_bits;
	kp.work_counter = d_work;
	// whole_words: every warp owns complete verdict words and writes them with plain stores (no zeroing, no atomics)
	kp.ppw_log2 = whole_words ? 5u : pick_ppw_log2(ctx, n, grid);
	const bool split = sb != nullptr && ctx->opt_split && ctx->link_model_h.size() == 1 && n >= (uint64_t)ctx->opt_split_min &&
	                   n < (1ull << 32) && !ctx->objects_h.empty();
	if (split) {
		if (ran_split) *ran_split = true;
		const uint64_t cap64 = std::min<uint64_t>(n * (uint64_t)ctx->opt_split_cap, 0xfffffff0ull);
		{
			const int rrc = reserve_split(ctx, *sb, n);
			if (rrc != MPTCC_OK) return rrc;
		}
		kp.cand_rec = sb->rec.p;
		kp.cand_rel = sb->rel.p;
		kp.cand_obj = sb->obj.p;
		kp.cand_pose = sb->pose.p;
		kp.cand_total = sb->ctl.p;
		kp.overflow = sb->ctl.p + 1;
		kp.work_counter2 = sb->ctl.p + 2;
		kp.cand_cap = (unsigned int)cap64;
		const bool np3 = ctx->opt_np3 != 0;
		const bool np2 = ctx->opt_np2 != 0 || np3;
		kp.cand_base = sb->base.p;
		kp.cand_cert = sb->cert.p;
		kp.xq = sb->xq.p;
		kp.xq_total = sb->ctl.p + 3;
		kp.xq_cap = (unsigned int)cap64;
		{
			// top of the agent tree staged in shared memory: whole sibling pairs only (odd count)
			const DevModel &am = ctx->agent_h.models[ctx->link_model_h[0]];
			uint32_t top = std::min<uint32_t>(2u * am.num_tris - 1u, (ctx->opt_qnode && ctx->env_q_ok && ctx->env_h.q_ok && ctx->agent_h.q_ok) ? kAgentTopMaxQ : kAgentTopMax);
			kp.qnode = (ctx->opt_qnode && ctx->env_q_ok && ctx->env_h.q_ok && ctx->agent_h.q_ok) ? 1u : 0u;
			// (both trees expanded per step: up to three pending pairs per level, 128 stack entries per lane)
			kp.expand4 = (ctx->opt_expand4 && 3u * std::max(ctx->env_h.max_depth, ctx->agent_h.max_depth) + 4u <= 2u * kLaneStackCap) ? 1u : 0u;
			if (ctx->opt_agent_top >= 0) top = std::min<uint32_t>(top, (uint32_t)ctx->opt_agent_top);
			if (np3) {
				// np3 stages agent nodes only on request (agent_top = n): measured on C5, the shared memory they take away
				// from L1 costs more than the LDS gathers save
				top = ctx->opt_agent_top > 0 ? std::min<uint32_t>(std::min<uint32_t>(2u * am.num_tris - 1u, np3_agent_top_max()), (uint32_t)ctx->opt_agent_top) : 0u;
				kp.expand4 = 0u;
			}
			if (top && top % 2 == 0) --top;
			kp.agent_top = top;
			kp.agent_swz = d.agent_swz.p;
		}
		CU(cudaMemsetAsync(sb->ctl.p, 0, 4 * sizeof(unsigned int), s));
		CU(cudaMemsetAsync(d_work, 0, sizeof(uint32_t), s));
		CU(cudaMemsetAsync(d_bits, 0, sizeof(uint32_t) * ((n + 31) / 32), s));
		const bool timed = ctx->opt_time_phases != 0;
		if (timed)
			for (auto &ev : d.phase_ev)
				if (!ev) CU(cudaEventCreate(&ev));
		if (timed) CU(cudaEventRecord(d.phase_ev[0], s));
		int e = launch_broadphase(kp, d.sm_count, ctx->count, np2, s);
		if (e != 0) return cuda_fail(ctx, (cudaError_t)e, "broadphase_kernel launch");
		if (timed) CU(cudaEventRecord(d.phase_ev[1], s));
		// ~1.4 candidates per pose-set on the benchmark scenes: long batches once every resident warp (2 CTAs x 8
		// warps per SM) can expect at least four of them
		const bool long_batches = ctx->opt_long_batches == 1 ||
		                          (ctx->opt_long_batches < 0 && n >= (uint64_t)d.sm_count * 16u * 128u * 3u);
		// Persisting-L2 window over the environment's traversal nodes while the narrowphase runs.  It pays when the
		// narrowphase dominates the launch and the nodes would otherwise compete for L2 (C5, blimp vs 73 MB of nodes:
		// 1.191 -> 1.156 ms per launch); with a 12-triangle agent on the same environment the launch is all broadphase and
		// the set-aside only takes L2 away from it (on-device edges: 11.7 -> 13.4 ms per 11 M pose-sets).  -1 = on for
		// agents of >= 128 triangles against >= 16 MB of environment nodes.
		const bool l2_persist = d.l2_persist_max > 0 &&
		                        (ctx->opt_l2_persist > 0 ||
		                         (ctx->opt_l2_persist < 0 && ctx->agent_h.models[ctx->link_model_h[0]].num_tris >= 128u &&
		                          (kp.qnode ? d.env.nodeq.n * sizeof(uint4) : d.env.node32.n * sizeof(float4)) >= (16u << 20)));
		if (l2_persist) {
			// keep (a fraction of) the environment's traversal nodes in the persisting part of L2 while the narrowphase runs
			cudaStreamAttrValue av;
			std::memset(&av, 0, sizeof av);
			const size_t bytes = kp.qnode ? d.env.nodeq.n * sizeof(uint4) : d.env.node32.n * sizeof(float4);
			av.accessPolicyWindow.base_ptr = kp.qnode ? (void *)d.env.nodeq.p : (void *)d.env.node32.p;
			av.accessPolicyWindow.num_bytes = std::min(bytes, d.l2_window_max);
			av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)d.l2_persist_max / (double)std::max<size_t>(1, av.accessPolicyWindow.num_bytes));
			av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
			av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
			CU(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &av));
		}
		if (np2) {
			e = np3 ? launch_narrowphase3(kp, d.sm_count, ctx->count, s) : launch_narrowphase2(kp, d.sm_count, ctx->count, long_batches, s);
			if (e != 0) return cuda_fail(ctx, (cudaError_t)e, np3 ? "narrowphase3_kernel launch" : "narrowphase2_kernel launch");
			if (timed) CU(cudaEventRecord(d.phase_ev[2], s));
			e = launch_exact(kp, d.sm_count, ctx->count, s);
			if (e != 0) return cuda_fail(ctx, (cudaError_t)e, "exact_kernel launch");
			ctx->ctr.n_kernel_launches++;
		} else {
			e = launch_narrowphase(kp, d.sm_count, ctx->count, long_batches, s);
			if (e != 0) return cuda_fail(ctx, (cudaError_t)e, "narrowphase_kernel launch");
			if (timed) CU(cudaEventRecord(d.phase_ev[2], s));
		}
		if (timed) CU(cudaEventRecord(d.phase_ev[3], s));
		if (l2_persist) {
			cudaStreamAttrValue av;
			std::memset(&av, 0, sizeof av);
			CU(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &av));  // num_bytes = 0: window off
		}
		// overflow fall-back: the fused kernel returns at once unless the candidate list overflowed
		CU(cudaMemsetAsync(d_work, 0, sizeof(uint32_t), s));
		kp.run_if = kp.overflow;
		kp.ppw_log2 = 4;  // atomicOr into the words the split kernels already wrote
		e = launch_collide(kp, grid, ctx->count, s);
		if (e != 0) return cuda_fail(ctx, (cudaError_t)e, "collide_kernel (fall-back) launch");
		if (timed) CU(cudaEventRecord(d.phase_ev[4], s));
		ctx->ctr.n_kernel_launches += 3;
		ctx->ctr.n_pose_sets += n;
		return enqueue_self(ctx, d, kp, d_work, grid, s);
	}
	const uint64_t chunks = (n + (1ull << kp.ppw_log2) - 1) >> kp.ppw_log2;
	int g = grid;
	const uint64_t need = (chunks + kWarpsPerCta - 1) / kWarpsPerCta;
	if ((uint64_t)g >= need) {
		g = (int)std::max<uint64_t>(need, 1);
		kp.work_counter = nullptr;  // every warp gets (at most) one chunk: no counter to reset
	} else {
		CU(cudaMemsetAsync(d_work, 0, sizeof(uint32_t), s));
	}
	if (kp.ppw_log2 < 5) CU(cudaMemsetAsync(d_bits, 0, sizeof(uint32_t) * ((n + 31) / 32), s));
	const int e = launch_collide(kp, g, ctx->count, s);
	if (e != 0) return cuda_fail(ctx, (cudaError_t)e, "collide_kernel launch");
	ctx->ctr.n_kernel_launches++;
	ctx->ctr.n_pose_sets += n;
	return enqueue_self(ctx, d, kp, d_work, grid, s);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
extern "C" {

int mptcc_create(mptcc_ctx **out, int n_devices, const int *device_ids) {
	if (!out || n_devices < 1) return MPTCC_ERR_INVALID;
	*out = nullptr;
	int avail = 0;
	cudaError_t e = cudaGetDeviceCount(&avail);
	if (e != cudaSuccess || avail < 1) {
		cudaGetLastError();
		return MPTCC_ERR_NO_DEVICE;
	}
	std::unique_ptr<mptcc_ctx> c(new mptcc_ctx);
	mptcc_ctx *ctx = c.get();
	ctx->devs.reserve((size_t)n_devices);
	// Every device is entered in ctx->devs BEFORE its resources are created, so that a failure half way releases
	// whatever exists (streams, events, pinned blocks, device buffers) through mptcc_destroy.
	auto bail = [&c](int code) {
		mptcc_destroy(c.release());
		return code;
	};
	for (int i = 0; i < n_devices; ++i) {
		const int id = device_ids ? device_ids[i] : i;
		if (id < 0 || id >= avail) return bail(MPTCC_ERR_NO_DEVICE);
		cudaDeviceProp prop;
		if (cudaGetDeviceProperties(&prop, id) != cudaSuccess) return bail(MPTCC_ERR_NO_DEVICE);
		if (prop.major != 10) return bail(MPTCC_ERR_NO_DEVICE);  // the kernels are built for sm_100a only
		ctx->devs.emplace_back();
		Device &d = ctx->devs.back();
		d.id = id;
		d.sm_count = prop.multiProcessorCount;
		d.l2_persist_max = (size_t)std::max(0, prop.persistingL2CacheMaxSize);
		d.l2_window_max = (size_t)std::max(0, prop.accessPolicyMaxWindowSize);
		if (cudaSetDevice(d.id) != cudaSuccess) return bail(MPTCC_ERR_CUDA);
		if (d.l2_persist_max > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, d.l2_persist_max) != cudaSuccess) {
			cudaGetLastError();
			d.l2_persist_max = 0;
		}
		for (int k = 0; k < kPipe; ++k) {
			if (cudaStreamCreateWithFlags(&d.stream[k], cudaStreamNonBlocking) != cudaSuccess) return bail(MPTCC_ERR_CUDA);
			if (cudaEventCreate(&d.ev0[k]) != cudaSuccess || cudaEventCreate(&d.ev1[k]) != cudaSuccess) return bail(MPTCC_ERR_CUDA);
			if (d.work[k].reserve(1) != cudaSuccess) return bail(MPTCC_ERR_CUDA);
		}
		for (int t = 0; t < kInFlight; ++t) {
			for (int k = 0; k < kPipe; ++k)
				if (cudaEventCreateWithFlags(&d.done[t][k], cudaEventDisableTiming) != cudaSuccess) return bail(MPTCC_ERR_CUDA);
			if (cudaMallocHost(reinterpret_cast<void **>(&d.ctl_h[t]), sizeof(unsigned int) * 4 * kPipe) != cudaSuccess)
				return bail(MPTCC_ERR_CUDA);
		}
		if (cudaHostAlloc(reinterpret_cast<void **>(&d.small_in), kSmallBytes, cudaHostAllocMapped) != cudaSuccess ||
		    cudaHostAlloc(reinterpret_cast<void **>(&d.small_out), kSmallBytes / 24 / 8 + 64, cudaHostAllocMapped) != cudaSuccess)
			return bail(MPTCC_ERR_CUDA);
		if (cudaHostGetDevicePointer(reinterpret_cast<void **>(&d.small_in_dev), d.small_in, 0) != cudaSuccess ||
		    cudaHostGetDevicePointer(reinterpret_cast<void **>(&d.small_out_dev), d.small_out, 0) != cudaSuccess)
			return bail(MPTCC_ERR_CUDA);
		if (d.work_user.reserve(1) != cudaSuccess) return bail(MPTCC_ERR_CUDA);
		if (d.counters.reserve(CTR_N) != cudaSuccess) return bail(MPTCC_ERR_CUDA);
		if (cudaMemset(d.counters.p, 0, sizeof(unsigned long long) * CTR_N) != cudaSuccess) return bail(MPTCC_ERR_CUDA);
	}
	*out = c.release();
	// MPTCC_OPTIONS="name=value,name=value": options applied to every new context (A/B runs of the test-suite and the
	// bench against a kernel option without touching their code); unknown names are ignored here
	if (const char *envo = std::getenv("MPTCC_OPTIONS")) {
		std::string all(envo);
		size_t pos = 0;
		while (pos < all.size()) {
			size_t end = all.find(',', pos);
			if (end == std::string::npos) end = all.size();
			const std::string kv = all.substr(pos, end - pos);
			const size_t eq = kv.find('=');
			if (eq != std::string::npos) mptcc_set_option(*out, kv.substr(0, eq).c_str(), std::atol(kv.c_str() + eq + 1));
			pos = end + 1;
		}
		(*out)->err.clear();
	}
	return MPTCC_OK;
}

void mptcc_destroy(mptcc_ctx *ctx) {
	if (!ctx) return;
	for (auto &d : ctx->devs) {
		cudaSetDevice(d.id);
		cudaDeviceSynchronize();
		d.env.release();
		d.agent.release();
		d.objects.release();
		d.obj_aabb.release();
		d.bp_nodes.release();
		d.link_model.release();
		d.self_objects.release();
		d.agent_swz.release();
		d.grid_cells.release();
		d.grid_entries.release();
		d.counters.release();
		d.work_user.release();
		for (auto &sb : d.split) sb.release();
		for (auto &ev : d.phase_ev)
			if (ev) cudaEventDestroy(ev);
		d.edge_in.release();
		d.edge_off.release();
		d.edge_poses.release();
		d.edge_bits.release();
		d.edge_safe.release();
		d.edge_work.release();
		if (d.small_in) cudaFreeHost(d.small_in);
		if (d.small_out) cudaFreeHost(d.small_out);
		for (int t = 0; t < kInFlight; ++t) {
			for (int k = 0; k < kPipe; ++k)
				if (d.done[t][k]) cudaEventDestroy(d.done[t][k]);
			if (d.ctl_h[t]) cudaFreeHost(d.ctl_h[t]);
		}
		for (int k = 0; k < kPipe; ++k) {
			d.poses[k].release();
			d.bits[k].release();
			d.work[k].release();
			if (d.stream[k]) cudaStreamDestroy(d.stream[k]);
			if (d.ev0[k]) cudaEventDestroy(d.ev0[k]);
			if (d.ev1[k]) cudaEventDestroy(d.ev1[k]);
		}
	}
	delete ctx;
}

const char *mptcc_last_error(const mptcc_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int mptcc_num_devices(const mptcc_ctx *ctx) { return ctx ? (int)ctx->devs.size() : 0; }

int mptcc_set_environment(mptcc_ctx *ctx, int n_objects, const mptcc_mesh *meshes, const double *obj_pose7) {
	if (!ctx || n_objects < 0 || (n_objects > 0 && (!meshes || !obj_pose7))) return fail(ctx, MPTCC_ERR_INVALID, "bad argument");
	if (n_objects >= (1 << 26)) return fail(ctx, MPTCC_ERR_LIMIT, "too many environment objects");
	{
		const int drc = drain(ctx);
		if (drc != MPTCC_OK) return drc;
	}
	ctx->env_h.clear();
	ctx->env_q_ok = true;
	ctx->objects_h.clear();
	ctx->obj_aabb_h.clear();
	ctx->bp_nodes_h.clear();
	ctx->env_set = false;
	// The OBB trees of the objects are independent: built by a few host threads, a block of objects at a time
	// (the flattening below stays sequential so that the packed order is the object order).
	constexpr int kBlock = 256;
	std::vector<HostBvh> block;
	std::vector<char> block_ok;
	for (int i = 0; i < n_objects; ++i) {
		if (i % kBlock == 0) {
			const int cnt = std::min(kBlock, n_objects - i);
			block.assign((size_t)cnt, HostBvh());
			block_ok.assign((size_t)cnt, 0);
			std::atomic<int> next(0);
			auto work = [&]() {
				for (int k = next.fetch_add(1); k < cnt; k = next.fetch_add(1)) {
					const mptcc_mesh &m = meshes[i + k];
					block_ok[(size_t)k] = build_bvh(m.xyz, m.nv, m.tri, m.nt, block[(size_t)k]) ? 1 : 0;
				}
			};
			const unsigned hw = std::thread::hardware_concurrency();
			const int n_threads = std::max(1, std::min<int>({(int)(hw ? hw : 1u), 16, cnt}));
			std::vector<std::thread> pool;
			for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
			work();
			for (auto &t : pool) t.join();
		}
		const HostBvh &b = block[(size_t)(i % kBlock)];
		if (b.num_tris() >= (1u << 22)) ctx->env_q_ok = false;
		if (!block_ok[(size_t)(i % kBlock)])
			return fail(ctx, MPTCC_ERR_INVALID, "environment object " + std::to_string(i) + " has an empty mesh or a triangle index beyond its vertices");
		DevObject o;
		std::memset(&o, 0, sizeof o);
		o.model = ctx->env_h.add(b);
		// fcl::Transform3f(Quaternion3f(w,x,y,z), Vec3f(x,y,z)) -- fcl_helpers::parseTransform
		const double *ps = obj_pose7 + 7 * (size_t)i;
		const double q[4] = {ps[3], ps[4], ps[5], ps[6]};
		o.T[0] = ps[0];
		o.T[1] = ps[1];
		o.T[2] = ps[2];
		quat_to_rotation(q, o.R);
		// fcl::CollisionObject::computeAABB
		if (q[0] == 1.0 && q[1] == 0.0 && q[2] == 0.0 && q[3] == 0.0) {
			for (int k = 0; k < 3; ++k) {
				o.aabb[k] = b.aabb_min[k] + o.T[k];
				o.aabb[3 + k] = b.aabb_max[k] + o.T[k];
			}
		} else {
			double c[3];
			quat_rotate(q, b.aabb_center, c);
			for (int k = 0; k < 3; ++k) {
				const double ck = c[k] + o.T[k];
				o.aabb[k] = ck - b.aabb_radius;
				o.aabb[3 + k] = ck + b.aabb_radius;
			}
		}
		{
			bool ident = o.T[0] == 0.0 && o.T[1] == 0.0 && o.T[2] == 0.0;
			for (int k = 0; k < 9; ++k) ident = ident && o.R[k] == ((k % 4 == 0) ? 1.0 : 0.0);
			o.identity = ident ? 1u : 0u;
		}
		ctx->objects_h.push_back(o);
	}
	sort_objects_morton(ctx->objects_h);
	{
		const size_t n = ctx->objects_h.size(), np = (n + kClusterSize - 1) / kClusterSize * kClusterSize;
		ctx->obj_aabb_h.assign(6 * np, 0.0);
		for (size_t i = 0; i < np; ++i)
			for (int k = 0; k < 6; ++k)  // padding: empty boxes (min = +max, max = -max) never overlap
				ctx->obj_aabb_h[k * np + i] = i < n ? ctx->objects_h[i].aabb[k] : (k < 3 ? DBL_MAX : -DBL_MAX);
	}
	ctx->bp_nodes_h.clear();
	if (!ctx->objects_h.empty()) {
		BpBuilder bp(ctx->objects_h);
		bp.emit(0, (ctx->objects_h.size() + kClusterSize - 1) / kClusterSize);
		ctx->bp_nodes_h.swap(bp.nodes);
		for (int k = 0; k < 3; ++k) {
			ctx->bp_origin[k] = bp.origin[k];
			ctx->bp_inv_cell[k] = bp.inv_cell[k];
		}
	}
	const int rc = upload_scene_part(ctx, true);
	if (rc != MPTCC_OK) return rc;
	ctx->env_set = true;
	if (ctx->agent_set) return build_grid(ctx);
	return MPTCC_OK;
}

int mptcc_set_agent(mptcc_ctx *ctx, int n_links, const mptcc_mesh *link_meshes) {
	if (!ctx || n_links < 1 || !link_meshes) return fail(ctx, MPTCC_ERR_INVALID, "bad argument");
	{
		const int drc = drain(ctx);
		if (drc != MPTCC_OK) return drc;
	}
	ctx->agent_h.clear();
	ctx->link_model_h.clear();
	ctx->agent_set = false;
	std::map<std::pair<const void *, const void *>, uint32_t> seen;
	for (int j = 0; j < n_links; ++j) {
		const mptcc_mesh &m = link_meshes[j];
		const auto key = std::make_pair((const void *)m.xyz, (const void *)m.tri);
		auto it = seen.find(key);
		if (it != seen.end()) {
			ctx->link_model_h.push_back(it->second);
			continue;
		}
		if (m.nt >= (1u << 21)) return fail(ctx, MPTCC_ERR_LIMIT, "agent link meshes are limited to 2^21 triangles");
		HostBvh b;
		if (!build_bvh(m.xyz, m.nv, m.tri, m.nt, b))
			return fail(ctx, MPTCC_ERR_INVALID, "agent link " + std::to_string(j) + " has an empty mesh");
		const uint32_t id = ctx->agent_h.add(b, true);  // breadth-first device ids: the top of the tree is one contiguous block
		seen[key] = id;
		ctx->link_model_h.push_back(id);
	}
	{
		// link model 0's traversal nodes in the order the np3 kernel keeps them in shared memory: part k (16 bytes) of
		// node i at position (k + (i >> 1)) & 3 of the node's 64 bytes (np3_kernels.cu, node_shared_swz)
		const DevModel &am = ctx->agent_h.models[ctx->link_model_h[0]];
		const uint32_t nn = 2u * am.num_tris - 1u;
		ctx->agent_swz_h.assign(4 * (size_t)nn, make_float4(0.f, 0.f, 0.f, 0.f));
		for (uint32_t i = 0; i < nn; ++i)
			for (uint32_t k = 0; k < 4; ++k)
				ctx->agent_swz_h[4 * (size_t)i + ((k + (i >> 1)) & 3u)] = ctx->agent_h.node32[4 * ((size_t)am.node_base + i) + k];
	}
	ctx->self_objects_h.assign(ctx->link_model_h.size(), DevObject{});
	for (size_t j = 0; j < ctx->link_model_h.size(); ++j) ctx->self_objects_h[j].model = ctx->link_model_h[j];
	const int rc = upload_scene_part(ctx, false);
	if (rc != MPTCC_OK) return rc;
	ctx->agent_set = true;
	if (ctx->env_set) return build_grid(ctx);
	return MPTCC_OK;
}

int mptcc_wait(mptcc_ctx *ctx, uint64_t ticket) {
	if (!ctx) return MPTCC_ERR_INVALID;
	if (ticket == 0) ticket = ctx->submitted;
	if (ticket > ctx->submitted) return fail(ctx, MPTCC_ERR_INVALID, "unknown ticket");
	bool overflowed = false;
	// the streams run the submits in order: waiting for a ticket completes every earlier one
	for (uint64_t t = ctx->completed + 1; t <= ticket; ++t) {
		const int slot = (int)((t - 1) % kInFlight);
		for (Device &d : ctx->devs) {
			if (d.used[slot] == 0) continue;
			CU(cudaSetDevice(d.id));
			for (size_t k = 0; k < d.used[slot]; ++k) {
				CU(cudaEventSynchronize(d.done[slot][k]));
				overflowed = overflowed || (d.ctl_h[slot][4 * k + 1] & 1u) != 0;  // [1] = overflow flag of the last split launch
				if (t == ctx->submitted) {
					float ms = 0.f;  // device time from the first kernel of the call (stream 0) to the end of the last kernel
					if (cudaEventElapsedTime(&ms, d.ev0[0], d.ev1[k]) == cudaSuccess) d.last_ms = std::max(d.last_ms, ms);
				}
			}
			d.used[slot] = 0;
		}
		ctx->completed = t;
	}
	// The fused fall-back already produced the right verdicts; a scene this dense gets a longer candidate list
	// from the next call on.
	if (overflowed && ctx->opt_split_cap < 64) ctx->opt_split_cap = std::min<long>(64, ctx->opt_split_cap * 2);
	return MPTCC_OK;
}

int mptcc_check_poses(mptcc_ctx *ctx, uint64_t n, const double *poses, int fmt, uint32_t *verdict_bits) {
	// Planner-sized calls (one edge, a few hundred pose-sets): no copies are enqueued at all -- the poses are put in
	// pinned host memory the fused kernel reads in place, the verdict words are stored straight into pinned
	// memory, and the only CUDA calls are one 4-byte memset, the launch and the synchronise.
	if (ctx && n > 0 && poses && verdict_bits && (fmt >= MPTCC_POSE_RT12 && fmt <= MPTCC_POSE_XYZCS) &&
	    ctx->submitted == ctx->completed && ctx->opt_ppw == 0 && !ctx->opt_self_collision && !ctx->devs.empty()) {
		const size_t bytes = n * pose_rec_len((unsigned)fmt) * ctx->link_model_h.size() * sizeof(double);
		if (bytes <= kSmallBytes && check_ready(ctx) == MPTCC_OK) {
			Device &d = ctx->devs[0];
			CU(cudaSetDevice(d.id));
			std::memcpy(d.small_in, poses, bytes);
			const int rc = enqueue_kernel(ctx, d, n, d.small_in_dev, fmt, d.small_out_dev, d.work[0].p, d.stream[0], nullptr, true);
			if (rc != MPTCC_OK) return rc;
			CU(cudaStreamSynchronize(d.stream[0]));
			std::memcpy(verdict_bits, d.small_out, sizeof(uint32_t) * ((n + 31) / 32));
			d.last_ms = 0.f;
			return MPTCC_OK;
		}
	}
	uint64_t ticket = 0;
	const int rc = mptcc_submit_poses(ctx, n, poses, fmt, verdict_bits, &ticket);
	if (rc != MPTCC_OK) return rc;
	return mptcc_wait(ctx, ticket);
}

int mptcc_submit_poses(mptcc_ctx *ctx, uint64_t n, const double *poses, int fmt, uint32_t *verdict_bits, uint64_t *ticket) {
	int rc = check_ready(ctx);
	if (rc != MPTCC_OK) return rc;
	if (!ticket) return fail(ctx, MPTCC_ERR_INVALID, "null ticket");
	if (fmt < MPTCC_POSE_RT12 || fmt > MPTCC_POSE_XYZCS) return fail(ctx, MPTCC_ERR_INVALID, "unknown pose format");
	*ticket = ctx->submitted;  // an empty batch completes with whatever is outstanding
	if (n == 0) return MPTCC_OK;
	if (!poses || !verdict_bits) return fail(ctx, MPTCC_ERR_INVALID, "null buffer");
	if (ctx->submitted - ctx->completed >= (uint64_t)kInFlight) {
		rc = mptcc_wait(ctx, ctx->completed + 1);
		if (rc != MPTCC_OK) return rc;
	}
	const bool idle_at_submit = ctx->submitted == ctx->completed;
	const int tslot = (int)(ctx->submitted % kInFlight);
	const size_t rec = pose_rec_len((unsigned)fmt) * ctx->link_model_h.size();  // doubles per pose-set
	const uint64_t n_words = (n + 31) / 32;
	const size_t nd = ctx->devs.size();
	// contiguous word-aligned slices per device, cut into sub-batches that ping-pong over two streams
	const uint64_t words_per_dev = (n_words + nd - 1) / nd;
	const uint64_t sub_words = std::max<uint64_t>((uint64_t)ctx->opt_sub_batch / 32, 1);
	struct Job {
		size_t dev;
		uint64_t w0, w1;
	};
	std::vector<std::vector<Job>> jobs(nd);
	size_t max_jobs = 0;
	for (size_t di = 0; di < nd; ++di) {
		const uint64_t b = std::min(n_words, di * words_per_dev), e = std::min(n_words, (di + 1) * words_per_dev);
		// split evenly into ceil(len/sub_words) pieces (ramping the first pieces up to shorten the one copy no
		// kernel can hide was measured slower: more, smaller launches)
		const uint64_t len = e - b;
		if (len == 0) continue;
		// (2^19 instead of 2^18 pose-sets per piece: the same blocking time on the benchmark batch, 8 % more with two calls in
		// flight -- short launches end in a tail of long candidates; two pieces at least, so that a copy hides behind a kernel)
		const uint64_t pieces = std::max<uint64_t>((len + sub_words - 1) / sub_words, len >= 2048 ? 2 : 1);
		const uint64_t per = (len + pieces - 1) / pieces;
		// (ramping the first pieces up was measured again with the round-2 kernels: 2.095 against 2.109 ms blocking, but
		// 2.0 against 1.7 ms with two calls in flight)
		uint64_t w = b;
		if (ctx->opt_first_half && idle_at_submit && per * 32 * rec * sizeof(double) >= (24u << 20)) {
			// Nothing is in flight, so the copy of the first piece is the one no kernel hides: cut that piece in two
			// (blocking call on the benchmark batch 2.13 -> 1.98 ms).  With an earlier submit still running the copy hides
			// behind its kernels and the extra, shorter launch only costs (two calls in flight: 1.49 -> 1.83 ms), so the
			// pieces stay whole then.  The extra launch costs ~0.08 ms, half a copy of 24 MB saves ~0.11 ms: smaller first
			// pieces stay whole as well (C1-C3 of BASELINE.md lose 3-6 % otherwise).
			// (an eighth, an eighth and a quarter of the call first: 2.00 ms, no better)
			const uint64_t h = per / 2;
			jobs[di].push_back(Job{di, w, w + h});
			jobs[di].push_back(Job{di, w + h, std::min(e, w + per)});
			w += per;
		}
		for (; w < e; w += per) jobs[di].push_back(Job{di, w, std::min(e, w + per)});
		max_jobs = std::max(max_jobs, jobs[di].size());
		ctx->devs[di].last_ms = 0.f;
	}
	// Size every staging set for the largest job up front: inside the loop nothing is reallocated, so a
	// staging set can be reused by the next job of ITS stream without any host-side wait (stream order protects
	// the buffers) and the host only blocks once, at the end.
	for (size_t di = 0; di < nd; ++di) {
		if (jobs[di].empty()) continue;
		Device &d = ctx->devs[di];
		uint64_t max_cnt = 0, max_words = 0;
		for (const Job &jb : jobs[di]) {
			max_cnt = std::max<uint64_t>(max_cnt, std::min<uint64_t>(n, jb.w1 * 32) - jb.w0 * 32);
			max_words = std::max<uint64_t>(max_words, jb.w1 - jb.w0);
		}
		CU(cudaSetDevice(d.id));
		for (size_t k = 0; k < std::min<size_t>(jobs[di].size(), kPipe); ++k) {
			CU(d.poses[k].reserve(max_cnt * rec));
			CU(d.bits[k].reserve(max_words));
			rc = reserve_split(ctx, d.split[k], max_cnt);
			if (rc != MPTCC_OK) return rc;
		}
	}
	// round-robin over devices so that every GPU's queue fills concurrently
	for (size_t ji = 0; ji < max_jobs; ++ji) {
		for (size_t di = 0; di < nd; ++di) {
			if (ji >= jobs[di].size()) continue;
			Device &d = ctx->devs[di];
			const Job &jb = jobs[di][ji];
			const int k = (int)(ji % kPipe);
			const uint64_t p0 = jb.w0 * 32, p1 = std::min<uint64_t>(n, jb.w1 * 32), cnt = p1 - p0;
			CU(cudaSetDevice(d.id));
			CU(cudaMemcpyAsync(d.poses[k].p, poses + p0 * rec, sizeof(double) * cnt * rec, cudaMemcpyHostToDevice, d.stream[k]));
			if (ji < (size_t)kPipe) CU(cudaEventRecord(d.ev0[k], d.stream[k]));  // first kernel of this stream
			bool ran_split = false;
			rc = enqueue_kernel(ctx, d, cnt, d.poses[k].p, fmt, d.bits[k].p, d.work[k].p, d.stream[k], &d.split[k], false, &ran_split);
			if (rc != MPTCC_OK) return rc;
			d.split_ran[k] = ran_split;  // (the last job of this stream counts: its flags are the ones read back)
			CU(cudaEventRecord(d.ev1[k], d.stream[k]));  // re-recorded by every job: the last one counts
			CU(cudaMemcpyAsync(verdict_bits + jb.w0, d.bits[k].p, sizeof(uint32_t) * (jb.w1 - jb.w0), cudaMemcpyDeviceToHost,
			                   d.stream[k]));
		}
	}
	// completion marker of this submit on both streams + the overflow flags of its last split launches
	for (size_t di = 0; di < nd; ++di) {
		Device &d = ctx->devs[di];
		d.used[tslot] = 0;
		if (jobs[di].empty()) continue;
		CU(cudaSetDevice(d.id));
		const size_t used = std::min<size_t>(jobs[di].size(), kPipe);
		for (size_t k = 0; k < used; ++k) {
			d.ctl_h[tslot][4 * k + 1] = 0;
			if (d.split[k].ctl.p && d.split_ran[k])  // (a stale flag of an earlier launch must not grow the list again)
				CU(cudaMemcpyAsync(d.ctl_h[tslot] + 4 * k, d.split[k].ctl.p, sizeof(unsigned int) * 4, cudaMemcpyDeviceToHost,
				                   d.stream[k]));
			CU(cudaEventRecord(d.done[tslot][k], d.stream[k]));
		}
		d.used[tslot] = used;
	}
	*ticket = ++ctx->submitted;
	return MPTCC_OK;
}

int mptcc_check_segments(mptcc_ctx *ctx, uint64_t n_segments, const uint64_t *seg_begin, const double *poses, int fmt,
                         uint8_t *segment_safe) {
	if (!ctx) return MPTCC_ERR_INVALID;
	if (n_segments == 0) return MPTCC_OK;
	if (!seg_begin || !segment_safe) return fail(ctx, MPTCC_ERR_INVALID, "null buffer");
	const uint64_t first = seg_begin[0], n = seg_begin[n_segments] - first;
	for (uint64_t s = 0; s < n_segments; ++s)
		if (seg_begin[s + 1] < seg_begin[s]) return fail(ctx, MPTCC_ERR_INVALID, "seg_begin must be non-decreasing");
	std::vector<uint32_t> bits((n + 31) / 32, 0u);
	const size_t rec = pose_rec_len((unsigned)fmt) * ctx->link_model_h.size();
	const int rc = mptcc_check_poses(ctx, n, poses ? poses + first * rec : nullptr, fmt, bits.data());
	if (rc != MPTCC_OK) return rc;
	for (uint64_t s = 0; s < n_segments; ++s) {
		uint8_t safe = 1;
		for (uint64_t i = seg_begin[s] - first; i < seg_begin[s + 1] - first; ++i)
			if ((bits[i >> 5] >> (i & 31)) & 1u) {
				safe = 0;
				break;
			}
		segment_safe[s] = safe;
	}
	return MPTCC_OK;
}

// Number of pose-sets Omnidirectional::getPoses produces for one edge (omnidirectional.hpp:216-244).
static uint64_t omni_pose_count(const double *e, double dt) {
	const double dx = e[3] - e[0], dy = e[4] - e[1], dz = e[5] - e[2];
	const double dist = std::sqrt(dx * dx + dy * dy + dz * dz);
	const double ratio = dist / dt;
	const unsigned int iterations = ratio >= 1.0 ? (ratio < 4294967295.0 ? (unsigned int)ratio : 4294967295u) : 0u;
	if (iterations < 1) return 2;
	return (uint64_t)iterations + (((double)iterations * dt < dist) ? 1u : 0u);
}

int mptcc_check_omni_edges(mptcc_ctx *ctx, uint64_t n_edges, const double *start_end, double dt, uint8_t *edge_safe) {
	int rc = check_ready(ctx);
	if (rc != MPTCC_OK) return rc;
	if (ctx->link_model_h.size() != 1) return fail(ctx, MPTCC_ERR_STATE, "mptcc_check_omni_edges needs a single-link agent");
	if (!(dt > 0.0)) return fail(ctx, MPTCC_ERR_INVALID, "dt must be positive");
	if (n_edges == 0) return MPTCC_OK;
	if (!start_end || !edge_safe) return fail(ctx, MPTCC_ERR_INVALID, "null buffer");
	// sub-batches of edges whose poses fit the device-side pose buffer, one per device and round: every device of the
	// context gets its sub-batch enqueued before any of them is waited for, so the devices work concurrently
	const uint64_t max_poses = 1ull << 21;
	const size_t nd = ctx->devs.size();
	std::vector<std::vector<unsigned long long> > offs(nd);  // per device: must stay alive until its stream is done
	uint64_t e0 = 0;
	while (e0 < n_edges) {
		size_t used = 0;
		for (size_t di = 0; di < nd && e0 < n_edges; ++di, ++used) {
			std::vector<unsigned long long> &off = offs[di];
			off.assign(1, 0ull);
			uint64_t e1 = e0;
			while (e1 < n_edges && (off.back() < max_poses || e1 == e0)) {
				off.push_back(off.back() + omni_pose_count(start_end + 6 * e1, dt));
				++e1;
			}
			const uint64_t ne = e1 - e0, np = off.back();
			Device &d = ctx->devs[di];
			CU(cudaSetDevice(d.id));
			cudaStream_t s = d.stream[0];
			CU(d.edge_in.reserve(6 * ne));
			CU(d.edge_off.reserve(ne + 1));
			CU(d.edge_poses.reserve(3 * np));
			CU(d.edge_bits.reserve((np + 31) / 32));
			CU(d.edge_safe.reserve(ne));
			CU(d.edge_work.reserve(1));
			CU(cudaMemcpyAsync(d.edge_in.p, start_end + 6 * e0, sizeof(double) * 6 * ne, cudaMemcpyHostToDevice, s));
			CU(cudaMemcpyAsync(d.edge_off.p, off.data(), sizeof(unsigned long long) * (ne + 1), cudaMemcpyHostToDevice, s));
			int e = launch_expand_omni_edges(d.edge_in.p, d.edge_off.p, ne, np, dt, d.edge_poses.p, s);
			if (e != 0) return cuda_fail(ctx, (cudaError_t)e, "expand_omni_edges launch");
			rc = enqueue_kernel(ctx, d, np, d.edge_poses.p, MPTCC_POSE_T3, d.edge_bits.p, d.edge_work.p, s, &d.split[kPipe + 1]);
			if (rc != MPTCC_OK) return rc;
			e = launch_reduce_segments(d.edge_bits.p, d.edge_off.p, ne, d.edge_safe.p, s);
			if (e != 0) return cuda_fail(ctx, (cudaError_t)e, "reduce_segments launch");
			CU(cudaMemcpyAsync(edge_safe + e0, d.edge_safe.p, ne, cudaMemcpyDeviceToHost, s));
			e0 = e1;
		}
		for (size_t di = 0; di < used; ++di) {  // the round's buffers (host `offs`, device edge_*) are reused by the next one
			Device &d = ctx->devs[di];
			CU(cudaSetDevice(d.id));
			CU(cudaStreamSynchronize(d.stream[0]));
			note_overflow(ctx, d, d.split[kPipe + 1]);
		}
	}
	return MPTCC_OK;
}

int mptcc_check_poses_device(mptcc_ctx *ctx, int slot, uint64_t n, const double *d_poses, int fmt, uint32_t *d_bits,
                             void *cuda_stream) {
	int rc = check_ready(ctx);
	if (rc != MPTCC_OK) return rc;
	if (slot < 0 || slot >= (int)ctx->devs.size()) return fail(ctx, MPTCC_ERR_INVALID, "bad device slot");
	if (fmt < MPTCC_POSE_RT12 || fmt > MPTCC_POSE_XYZCS) return fail(ctx, MPTCC_ERR_INVALID, "unknown pose format");
	if (n == 0) return MPTCC_OK;
	if (!d_poses || !d_bits) return fail(ctx, MPTCC_ERR_INVALID, "null buffer");
	Device &d = ctx->devs[slot];
	CU(cudaSetDevice(d.id));
	cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : d.stream[0];
	return enqueue_kernel(ctx, d, n, d_poses, fmt, d_bits, d.work_user.p, s, &d.split[kPipe]);
}

int mptcc_synchronize(mptcc_ctx *ctx) {
	if (!ctx) return MPTCC_ERR_INVALID;
	for (auto &d : ctx->devs) {
		CU(cudaSetDevice(d.id));
		CU(cudaDeviceSynchronize());
		note_overflow(ctx, d, d.split[kPipe]);  // candidate list of mptcc_check_poses_device
	}
	return MPTCC_OK;
}

void *mptcc_alloc_pinned(size_t bytes) {
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	return p;
}

void mptcc_free_pinned(void *p) {
	if (p) cudaFreeHost(p);
}

int mptcc_enable_counters(mptcc_ctx *ctx, int enable) {
	if (!ctx) return MPTCC_ERR_INVALID;
	ctx->count = enable != 0;
	return MPTCC_OK;
}

int mptcc_get_counters(mptcc_ctx *ctx, mptcc_counters *out) {
	if (!ctx || !out) return MPTCC_ERR_INVALID;
	unsigned long long sum[CTR_N] = {0};
	for (auto &d : ctx->devs) {
		unsigned long long h[CTR_N];
		CU(cudaSetDevice(d.id));
		CU(cudaDeviceSynchronize());
		CU(cudaMemcpy(h, d.counters.p, sizeof h, cudaMemcpyDeviceToHost));
		for (int i = 0; i < CTR_N; ++i) sum[i] += h[i];
	}
	*out = ctx->ctr;
	out->n_candidates = sum[CTR_CAND];
	out->n_bv_tests = sum[CTR_BV];
	out->n_leaf_tests = sum[CTR_LEAF];
	out->n_exact_tri = sum[CTR_EXACT];
	out->n_path_checks = sum[CTR_PATH];
	return MPTCC_OK;
}

int mptcc_reset_counters(mptcc_ctx *ctx) {
	if (!ctx) return MPTCC_ERR_INVALID;
	for (auto &d : ctx->devs) {
		CU(cudaSetDevice(d.id));
		CU(cudaDeviceSynchronize());
		CU(cudaMemset(d.counters.p, 0, sizeof(unsigned long long) * CTR_N));
	}
	ctx->ctr = mptcc_counters{};
	return MPTCC_OK;
}

int mptcc_last_kernel_ms(mptcc_ctx *ctx, int slot, float *ms) {
	if (!ctx || !ms || slot < 0 || slot >= (int)ctx->devs.size()) return MPTCC_ERR_INVALID;
	*ms = ctx->devs[slot].last_ms;
	return MPTCC_OK;
}

int mptcc_last_phase_ms(mptcc_ctx *ctx, int slot, float *ms4) {
	if (!ctx || !ms4 || slot < 0 || slot >= (int)ctx->devs.size()) return MPTCC_ERR_INVALID;
	Device &d = ctx->devs[slot];
	if (!d.phase_ev[0]) return fail(ctx, MPTCC_ERR_STATE, "no timed split-pipeline launch yet (option time_phases)");
	CU(cudaSetDevice(d.id));
	CU(cudaEventSynchronize(d.phase_ev[4]));
	for (int k = 0; k < 4; ++k) CU(cudaEventElapsedTime(ms4 + k, d.phase_ev[k], d.phase_ev[k + 1]));
	return MPTCC_OK;
}

int mptcc_get_info(mptcc_ctx *ctx, const char *name, long *value) {
	if (!ctx || !name || !value) return MPTCC_ERR_INVALID;
	const std::string k(name);
	if (k == "qnode_active") *value = (ctx->opt_qnode && ctx->env_q_ok && ctx->env_h.q_ok && ctx->agent_h.q_ok) ? 1 : 0;
	else if (k == "grid_cells") *value = ctx->opt_grid ? (long)ctx->grid_cells_h.size() : 0;
	else if (k == "grid_entries") *value = ctx->opt_grid ? (long)ctx->grid_entries_h.size() : 0;
	else if (k == "agent_nodes") *value = (long)(ctx->agent_h.node32.size() / 4);
	else if (k == "env_nodes") *value = (long)(ctx->env_h.node32.size() / 4);
	else return fail(ctx, MPTCC_ERR_INVALID, "unknown info: " + k);
	return MPTCC_OK;
}

int mptcc_set_option(mptcc_ctx *ctx, const char *name, long value) {
	if (!ctx || !name) return MPTCC_ERR_INVALID;
	const std::string k(name);
	if (k == "poses_per_warp") {
		if (value < 0 || value > 64 || (value & (value - 1))) return fail(ctx, MPTCC_ERR_INVALID, "poses_per_warp: 0 or a power of two <= 64");
		ctx->opt_ppw = value;
	} else if (k == "ctas_per_sm") {
		if (value < 0 || value > 32) return fail(ctx, MPTCC_ERR_INVALID, "ctas_per_sm: 0..32");
		ctx->opt_ctas_per_sm = value;
	} else if (k == "long_batches") {
		ctx->opt_long_batches = value < 0 ? -1 : (value != 0);
	} else if (k == "self_collision") {
		if (ctx->opt_self_collision != (value != 0) && ctx->submitted != ctx->completed) {
			const int drc = mptcc_wait(ctx, 0);  // submits in flight keep the behaviour they were enqueued with
			if (drc != MPTCC_OK) return drc;
		}
		ctx->opt_self_collision = value != 0;
	} else if (k == "first_half") {
		ctx->opt_first_half = value != 0;
	} else if (k == "grid") {
		ctx->opt_grid = value != 0;
	} else if (k == "np3") {
		ctx->opt_np3 = value != 0;
	} else if (k == "np2") {
		ctx->opt_np2 = value != 0;
		if (value != 0) ctx->opt_np3 = 0;  // (np3 is the default; asking for np2 selects np2)
	} else if (k == "expand4") {
		ctx->opt_expand4 = value != 0;
	} else if (k == "l2_persist") {
		ctx->opt_l2_persist = value < 0 ? -1 : (value != 0);
	} else if (k == "qnode") {
		ctx->opt_qnode = value != 0;
	} else if (k == "agent_top") {
		if (value < -1) return fail(ctx, MPTCC_ERR_INVALID, "agent_top: -1 (auto) or a node count");
		ctx->opt_agent_top = value;
	} else if (k == "near_first") {
		ctx->opt_near_first = value != 0;
	} else if (k == "steal_bottom") {
		ctx->opt_steal_bottom = value != 0;
	} else if (k == "no_steal") {
		ctx->opt_no_steal = value != 0;
	} else if (k == "time_phases") {
		ctx->opt_time_phases = value != 0;
	} else if (k == "split") {
		ctx->opt_split = value != 0;
	} else if (k == "split_min") {
		if (value < 32) return fail(ctx, MPTCC_ERR_INVALID, "split_min: >= 32");
		ctx->opt_split_min = value;
	} else if (k == "split_cap") {
		if (value < 1 || value > 64) return fail(ctx, MPTCC_ERR_INVALID, "split_cap: 1..64 candidates per pose-set");
		ctx->opt_split_cap = value;
	} else if (k == "always_replay") {
		ctx->opt_always_replay = value != 0;
	} else if (k == "sub_batch") {
		if (value < 32) return fail(ctx, MPTCC_ERR_INVALID, "sub_batch: >= 32");
		ctx->opt_sub_batch = value;
	} else {
		return fail(ctx, MPTCC_ERR_INVALID, "unknown option " + k);
	}
	return MPTCC_OK;
}

// ---- host-only scene introspection (tests): a context without devices and the broadphase structures' answers ------
int mptcc_debug_create_host_only(mptcc_ctx **out) {
	if (!out) return MPTCC_ERR_INVALID;
	*out = new mptcc_ctx;  // no Device: set_environment / set_agent only build the host side, every check fails with NO_DEVICE
	return MPTCC_OK;
}

static int overlapping_of(const mptcc_ctx *ctx, const double *box6, const std::vector<uint32_t> &listed, double *aabbs, uint32_t cap) {
	int n = 0;
	for (uint32_t oi : listed) {
		const DevObject &o = ctx->objects_h[oi];
		bool ov = true;  // fcl::AABB::overlap: closed intervals
		for (int k = 0; k < 3; ++k) ov = ov && !(o.aabb[k] > box6[3 + k]) && !(o.aabb[3 + k] < box6[k]);
		if (!ov) continue;
		if ((uint32_t)n < cap && aabbs) std::memcpy(aabbs + 6 * (size_t)n, o.aabb, 48);
		++n;
	}
	return n;
}

int mptcc_debug_grid_lookup(mptcc_ctx *ctx, const double *box6, double *aabbs, uint32_t cap, uint32_t *n_listed) {
	if (!ctx || !box6) return MPTCC_ERR_INVALID;
	if (n_listed) *n_listed = 0;
	if (ctx->grid_cells_h.empty()) return -1;
	// the look-up of broadphase_grid_kernel (split_kernels.cu): the cell of the box's minimum corner
	int c[3];
	for (int k = 0; k < 3; ++k) {
		const double lo = box6[k], hi = box6[3 + k];
		const int a = grid_cell(lo, ctx->grid_origin[k], ctx->grid_inv[k], ctx->grid_n[k]);
		const int b = grid_cell(hi, ctx->grid_origin[k], ctx->grid_inv[k], ctx->grid_n[k]);
		if (!(std::fabs(lo) <= 1.7e308) || !(std::fabs(hi) <= 1.7e308) || b - a > 1 || b < a) return -2;
		c[k] = a;
	}
	const uint2 hdr = ctx->grid_cells_h[((size_t)c[0] * ctx->grid_n[1] + c[1]) * ctx->grid_n[2] + c[2]];
	std::vector<uint32_t> listed;
	for (uint32_t e = 0; e < hdr.y; ++e) listed.push_back((uint32_t)ctx->grid_entries_h[hdr.x + e].obj);
	if (n_listed) *n_listed = hdr.y;
	return overlapping_of(ctx, box6, listed, aabbs, cap);
}

int mptcc_debug_tree_lookup(mptcc_ctx *ctx, const double *box6, double *aabbs, uint32_t cap, uint32_t *n_listed) {
	if (!ctx || !box6) return MPTCC_ERR_INVALID;
	if (n_listed) *n_listed = 0;
	if (ctx->bp_nodes_h.empty()) return -1;
	// the walk of the fused kernel / broadphase_kernel (collide_device.cuh: pose_world_box's outward-rounded fp32 box,
	// quantize_pose_box, bp_tree_step)
	int qmin[3], qmax[3];
	for (int k = 0; k < 3; ++k) {
		float amin = (float)box6[k], amax = (float)box6[3 + k];
		if ((double)amin > box6[k]) amin = std::nextafterf(amin, -INFINITY);
		if ((double)amax < box6[3 + k]) amax = std::nextafterf(amax, INFINITY);
		const float lo = (amin - ctx->bp_origin[k]) * ctx->bp_inv_cell[k], hi = (amax - ctx->bp_origin[k]) * ctx->bp_inv_cell[k];
		qmin[k] = (int)std::floor(std::fmin(std::fmax(lo, -2.0f), 65538.0f)) - 1;
		qmax[k] = (int)std::ceil(std::fmax(std::fmin(hi, 65538.0f), -2.0f)) + 1;
	}
	std::vector<uint32_t> listed;
	const uint32_t nn = (uint32_t)ctx->bp_nodes_h.size();
	for (uint32_t node = 0; node < nn;) {
		const uint4 nd = ctx->bp_nodes_h[node];
		const int lx = (int)(nd.x & 0xffffu), ly = (int)(nd.x >> 16), lz = (int)(nd.y & 0xffffu);
		const int hx = (int)(nd.y >> 16), hy = (int)(nd.z & 0xffffu), hz = (int)(nd.z >> 16);
		const bool ov = !(lx > qmax[0] || ly > qmax[1] || lz > qmax[2] || hx < qmin[0] || hy < qmin[1] || hz < qmin[2]);
		const bool leaf = (nd.w >> 31) != 0u;
		if (ov && leaf) {
			const uint32_t cl = nd.w & 0x7fffffffu;
			for (uint32_t o = cl * kClusterSize; o < std::min<size_t>(ctx->objects_h.size(), (size_t)(cl + 1) * kClusterSize); ++o) listed.push_back(o);
		}
		node = (leaf || ov) ? node + 1u : nd.w;
	}
	if (n_listed) *n_listed = (uint32_t)listed.size();
	return overlapping_of(ctx, box6, listed, aabbs, cap);
}

// ---- host-only BVH introspection -------------------------------------------------------------------
int mptcc_bvh_build(const mptcc_mesh *mesh, mptcc_bvh **out) {
	if (!mesh || !out) return MPTCC_ERR_INVALID;
	std::unique_ptr<mptcc_bvh> b(new mptcc_bvh);
	if (!build_bvh(mesh->xyz, mesh->nv, mesh->tri, mesh->nt, b->b)) return MPTCC_ERR_INVALID;
	*out = b.release();
	return MPTCC_OK;
}

void mptcc_bvh_free(mptcc_bvh *b) { delete b; }
uint32_t mptcc_bvh_num_nodes(const mptcc_bvh *b) { return b ? b->b.num_nodes() : 0; }
uint32_t mptcc_bvh_depth(const mptcc_bvh *b) { return b ? b->b.depth : 0; }

void mptcc_bvh_get_nodes(const mptcc_bvh *b, double *axis, double *To, double *extent, int32_t *first_child,
                         int32_t *first_primitive, int32_t *num_primitives) {
	const HostBvh &h = b->b;
	const size_t n = h.num_nodes();
	if (axis) std::memcpy(axis, h.axis.data(), sizeof(double) * 9 * n);
	if (To) std::memcpy(To, h.center.data(), sizeof(double) * 3 * n);
	if (extent) std::memcpy(extent, h.extent.data(), sizeof(double) * 3 * n);
	if (first_child) std::memcpy(first_child, h.first_child.data(), sizeof(int32_t) * n);
	if (first_primitive) std::memcpy(first_primitive, h.first_primitive.data(), sizeof(int32_t) * n);
	if (num_primitives) std::memcpy(num_primitives, h.num_primitives.data(), sizeof(int32_t) * n);
}

int mptcc_bvh_get_quantised_nodes(const mptcc_bvh *b, int breadth_first, uint32_t *words, uint32_t *device_id, double *ctr3) {
	if (!b || !words) return -1;
	HostPack pk;
	pk.add(b->b, breadth_first != 0);
	const DevModel &m = pk.models[0];
	const size_t n = b->b.num_nodes();
	std::memcpy(words, pk.nodeq.data() + 2 * (size_t)m.node_base, sizeof(uint4) * 2 * n);
	if (ctr3) std::memcpy(ctr3, m.ctr, 24);
	if (device_id) {
		// the numbering HostPack::add uses (FCL's allocation order, or breadth-first with siblings adjacent)
		std::vector<uint32_t> order(n);
		if (breadth_first) {
			uint32_t head = 0, tail = 0;
			order[tail++] = 0;
			while (head < tail) {
				const uint32_t v = order[head++];
				if (b->b.first_child[v] >= 0) {
					order[tail++] = (uint32_t)b->b.first_child[v];
					order[tail++] = (uint32_t)b->b.first_child[v] + 1u;
				}
			}
		} else {
			for (uint32_t i = 0; i < n; ++i) order[i] = i;
		}
		for (uint32_t i = 0; i < n; ++i) device_id[order[i]] = i;
	}
	return pk.q_ok ? 1 : 0;
}

int mptcc_bvh_get_exact_arrays(const mptcc_bvh *b, int breadth_first, double *node64, uint32_t *anc, uint32_t *depth) {
	if (!b || !depth) return -1;
	HostPack pk;
	pk.add(b->b, breadth_first != 0);
	const DevModel &m = pk.models[0];
	*depth = m.depth;
	if (node64) std::memcpy(node64, pk.node64.data() + 16 * (size_t)m.node_base, sizeof(double) * 16 * b->b.num_nodes());
	if (anc) std::memcpy(anc, pk.anc.data() + m.anc_base, sizeof(uint32_t) * (size_t)m.num_tris * m.depth);
	return 0;
}

int mptcc_bvh_get_certificate_data(const mptcc_bvh *b, float *mu_chain, double *model6) {
	if (!b || !model6) return -1;
	HostPack pk;
	pk.add(b->b, false);
	const DevModel &m = pk.models[0];
	if (mu_chain) std::memcpy(mu_chain, pk.mu_chain.data(), sizeof(float) * m.num_tris);
	model6[0] = m.ctr[0];
	model6[1] = m.ctr[1];
	model6[2] = m.ctr[2];
	model6[3] = m.cmax1;
	model6[4] = m.esum;
	model6[5] = (double)m.r2;
	return 0;
}

void mptcc_bvh_get_primitive_indices(const mptcc_bvh *b, uint32_t *out) {
	std::memcpy(out, b->b.primitive_indices.data(), sizeof(uint32_t) * b->b.primitive_indices.size());
}

void mptcc_bvh_get_local_aabb(const mptcc_bvh *b, double *aabb6, double *center3, double *radius) {
	const HostBvh &h = b->b;
	if (aabb6) {
		std::memcpy(aabb6, h.aabb_min, 24);
		std::memcpy(aabb6 + 3, h.aabb_max, 24);
	}
	if (center3) std::memcpy(center3, h.aabb_center, 24);
	if (radius) *radius = h.aabb_radius;
}

}  // extern "C"
