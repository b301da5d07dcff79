// fcl_types.hpp -- stand-ins for the four FCL math types the toolkit's workspace API exposes
// (fcl::Vec3f, fcl::Matrix3f, fcl::Quaternion3f, fcl::Transform3f), for builds without FCL.
//
// The reference passes poses around as fcl::Transform3f (e.g. Map3D::safePoses,
// /root/reference/workspaces/map3d.hpp:39; Agent::getPoses, /root/reference/agents/omnidirectional.hpp:202).
// These classes only CARRY a pose to the C ABI (include/mptcc.h); all collision arithmetic happens in
// libmptcc.  What matters for parity is which constructor built the transform, because FCL keeps the
// rotation the caller gave (matrix or quaternion) and derives the other one [FCL-recalled C.1]:
//   Transform3f(T), Transform3f(R, T)  -> pose format MPTCC_POSE_RT12
//   Transform3f(q, T)                  -> pose format MPTCC_POSE_TQ7
// A build that does have FCL can keep its own fcl:: types and use the two pack_* functions of
// meshhandler.hpp as the only adapter.
#pragma once

#include <array>
#include <cstddef>

namespace fcl {

typedef double FCL_REAL;

class Vec3f {
public:
	Vec3f() : v{0, 0, 0} {}
	Vec3f(FCL_REAL x, FCL_REAL y, FCL_REAL z) : v{x, y, z} {}
	FCL_REAL operator[](std::size_t i) const { return v[i]; }
	FCL_REAL &operator[](std::size_t i) { return v[i]; }
	void setValue(FCL_REAL x, FCL_REAL y, FCL_REAL z) { v = {x, y, z}; }

private:
	std::array<FCL_REAL, 3> v;
};

class Matrix3f {
public:
	Matrix3f() { setIdentity(); }
	Matrix3f(FCL_REAL xx, FCL_REAL xy, FCL_REAL xz, FCL_REAL yx, FCL_REAL yy, FCL_REAL yz, FCL_REAL zx, FCL_REAL zy,
	         FCL_REAL zz)
	    : m{xx, xy, xz, yx, yy, yz, zx, zy, zz} {}
	void setIdentity() { m = {1, 0, 0, 0, 1, 0, 0, 0, 1}; }
	FCL_REAL operator()(std::size_t r, std::size_t c) const { return m[3 * r + c]; }
	FCL_REAL &operator()(std::size_t r, std::size_t c) { return m[3 * r + c]; }
	const FCL_REAL *data() const { return m.data(); }

private:
	std::array<FCL_REAL, 9> m;  // row-major
};

class Quaternion3f {
public:
	Quaternion3f() : q{1, 0, 0, 0} {}
	Quaternion3f(FCL_REAL w, FCL_REAL x, FCL_REAL y, FCL_REAL z) : q{w, x, y, z} {}  // (a,b,c,d) = (w,x,y,z)
	FCL_REAL getW() const { return q[0]; }
	FCL_REAL getX() const { return q[1]; }
	FCL_REAL getY() const { return q[2]; }
	FCL_REAL getZ() const { return q[3]; }

private:
	std::array<FCL_REAL, 4> q;
};

class Transform3f {
public:
	Transform3f() : from_quat(false) {}
	explicit Transform3f(const Vec3f &T_) : T(T_), from_quat(false) {}
	Transform3f(const Matrix3f &R_, const Vec3f &T_) : R(R_), T(T_), from_quat(false) {}
	Transform3f(const Quaternion3f &q_, const Vec3f &T_) : T(T_), q(q_), from_quat(true) {}

	const Vec3f &getTranslation() const { return T; }
	// Valid for transforms built from a matrix / a translation.  For quaternion-built transforms FCL derives
	// the matrix lazily (Quaternion3f::toRotation); libmptcc does the same on the device.
	const Matrix3f &getRotation() const { return R; }
	const Quaternion3f &getQuatRotation() const { return q; }
	bool builtFromQuaternion() const { return from_quat; }

private:
	Matrix3f R;
	Vec3f T;
	Quaternion3f q;
	bool from_quat;
};

}  // namespace fcl
