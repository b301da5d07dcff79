// fcl_types.hpp -- stand-ins for the four FCL math types the toolkit's workspace API exposes
// (fcl::Vec3f, fcl::Matrix3f, fcl::Quaternion3f, fcl::Transform3f), for builds without FCL.
//
// The reference passes poses around as fcl::Transform3f (e.g. Map3D::safePoses,
// /root/reference/workspaces/map3d.hpp:39; Agent::getPoses, /root/reference/agents/omnidirectional.hpp:202) and its
// agents and discretizations do a little arithmetic on them (agents/snake_trailers.hpp:411-459: Matrix3f * Matrix3f;
// discretizations/workspace/prmlite.hpp:184-211: Vec3f +, -, /, dot, Transform3f::setTransform).  Everything those
// files use is here, with FCL 0.3.x's semantics [FCL-recalled C.1]:
//   * Transform3f keeps the rotation the caller gave -- matrix (matrix_set) or quaternion -- and derives the other one
//     with Quaternion3f::fromRotation / toRotation, so getRotation() and getQuatRotation() are both always valid;
//   * which one was GIVEN decides the pose record format sent to libmptcc (include/mptcc.h), because FCL computes
//     with the given one:  Transform3f(T), Transform3f(R, T), setTransform(R, T)  -> RT12 (T3 / XYZCS when R is the
//     identity / a planar rotation);  Transform3f(q, T), setTransform(q, T), setQuatRotation -> TQ7.
// All collision arithmetic happens in libmptcc; these classes carry poses.  A build that does have FCL can keep its
// own fcl:: types and use MeshHandler::packPoses (meshhandler.hpp) as the only adapter.
#pragma once

#include <array>
#include <cmath>
#include <cstddef>

namespace fcl {

typedef double FCL_REAL;

class Vec3f {
public:
	Vec3f() : v{{0, 0, 0}} {}
	Vec3f(FCL_REAL x, FCL_REAL y, FCL_REAL z) : v{{x, y, z}} {}
	FCL_REAL operator[](std::size_t i) const { return v[i]; }
	FCL_REAL &operator[](std::size_t i) { return v[i]; }
	void setValue(FCL_REAL x, FCL_REAL y, FCL_REAL z) { v = {{x, y, z}}; }
	Vec3f operator+(const Vec3f &o) const { return Vec3f(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
	Vec3f operator-(const Vec3f &o) const { return Vec3f(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
	Vec3f operator-() const { return Vec3f(-v[0], -v[1], -v[2]); }
	Vec3f operator*(FCL_REAL t) const { return Vec3f(v[0] * t, v[1] * t, v[2] * t); }
	Vec3f operator/(FCL_REAL t) const { return Vec3f(v[0] / t, v[1] / t, v[2] / t); }
	Vec3f &operator+=(const Vec3f &o) { return *this = *this + o; }
	Vec3f &operator-=(const Vec3f &o) { return *this = *this - o; }
	Vec3f &operator*=(FCL_REAL t) { return *this = *this * t; }
	FCL_REAL dot(const Vec3f &o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
	Vec3f cross(const Vec3f &o) const {
		return Vec3f(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
	}
	FCL_REAL sqrLength() const { return dot(*this); }
	FCL_REAL length() const { return std::sqrt(sqrLength()); }
	Vec3f &normalize() {
		const FCL_REAL s = sqrLength();
		if (s > 0) *this = *this * (1.0 / std::sqrt(s));
		return *this;
	}
	bool operator==(const Vec3f &o) const { return v == o.v; }
	bool operator!=(const Vec3f &o) const { return !(*this == o); }

private:
	std::array<FCL_REAL, 3> v;
};

class Matrix3f {
public:
	Matrix3f() { setIdentity(); }
	Matrix3f(FCL_REAL xx, FCL_REAL xy, FCL_REAL xz, FCL_REAL yx, FCL_REAL yy, FCL_REAL yz, FCL_REAL zx, FCL_REAL zy,
	         FCL_REAL zz)
	    : m{{xx, xy, xz, yx, yy, yz, zx, zy, zz}} {}
	void setIdentity() { m = {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
	void setZero() { m.fill(0); }
	FCL_REAL operator()(std::size_t r, std::size_t c) const { return m[3 * r + c]; }
	FCL_REAL &operator()(std::size_t r, std::size_t c) { return m[3 * r + c]; }
	const FCL_REAL *data() const { return m.data(); }
	Vec3f getRow(std::size_t r) const { return Vec3f(m[3 * r], m[3 * r + 1], m[3 * r + 2]); }
	Vec3f getColumn(std::size_t c) const { return Vec3f(m[c], m[3 + c], m[6 + c]); }
	Vec3f operator*(const Vec3f &x) const { return Vec3f(getRow(0).dot(x), getRow(1).dot(x), getRow(2).dot(x)); }
	Matrix3f operator*(const Matrix3f &o) const {
		Matrix3f r;
		for (std::size_t i = 0; i < 3; ++i)
			for (std::size_t j = 0; j < 3; ++j) r(i, j) = getRow(i).dot(o.getColumn(j));
		return r;
	}
	Matrix3f &operator*=(const Matrix3f &o) { return *this = *this * o; }
	Matrix3f transpose() const { return Matrix3f(m[0], m[3], m[6], m[1], m[4], m[7], m[2], m[5], m[8]); }
	bool operator==(const Matrix3f &o) const { return m == o.m; }

private:
	std::array<FCL_REAL, 9> m;  // row-major
};

class Quaternion3f {
public:
	Quaternion3f() : q{{1, 0, 0, 0}} {}
	Quaternion3f(FCL_REAL w, FCL_REAL x, FCL_REAL y, FCL_REAL z) : q{{w, x, y, z}} {}  // (a,b,c,d) = (w,x,y,z)
	FCL_REAL getW() const { return q[0]; }
	FCL_REAL getX() const { return q[1]; }
	FCL_REAL getY() const { return q[2]; }
	FCL_REAL getZ() const { return q[3]; }
	bool isIdentity() const { return q[0] == 1 && q[1] == 0 && q[2] == 0 && q[3] == 0; }
	FCL_REAL dot(const Quaternion3f &o) const { return q[0] * o.q[0] + q[1] * o.q[1] + q[2] * o.q[2] + q[3] * o.q[3]; }
	Quaternion3f operator+(const Quaternion3f &o) const {
		return Quaternion3f(q[0] + o.q[0], q[1] + o.q[1], q[2] + o.q[2], q[3] + o.q[3]);
	}
	Quaternion3f operator*(FCL_REAL t) const { return Quaternion3f(q[0] * t, q[1] * t, q[2] * t, q[3] * t); }
	Quaternion3f operator*(const Quaternion3f &o) const {
		return Quaternion3f(q[0] * o.q[0] - q[1] * o.q[1] - q[2] * o.q[2] - q[3] * o.q[3],
		                    q[0] * o.q[1] + q[1] * o.q[0] + q[2] * o.q[3] - q[3] * o.q[2],
		                    q[0] * o.q[2] - q[1] * o.q[3] + q[2] * o.q[0] + q[3] * o.q[1],
		                    q[0] * o.q[3] + q[1] * o.q[2] - q[2] * o.q[1] + q[3] * o.q[0]);
	}
	Quaternion3f conj() const { return Quaternion3f(q[0], -q[1], -q[2], -q[3]); }
	Vec3f transform(const Vec3f &v) const {
		const Quaternion3f r = (*this) * Quaternion3f(0, v[0], v[1], v[2]) * conj();
		return Vec3f(r.q[1], r.q[2], r.q[3]);
	}
	// [FCL-recalled C.1] Quaternion3f::toRotation
	void toRotation(Matrix3f &R) const {
		const FCL_REAL twoX = 2.0 * q[1], twoY = 2.0 * q[2], twoZ = 2.0 * q[3];
		const FCL_REAL twoWX = twoX * q[0], twoWY = twoY * q[0], twoWZ = twoZ * q[0];
		const FCL_REAL twoXX = twoX * q[1], twoXY = twoY * q[1], twoXZ = twoZ * q[1];
		const FCL_REAL twoYY = twoY * q[2], twoYZ = twoZ * q[2], twoZZ = twoZ * q[3];
		R = Matrix3f(1.0 - (twoYY + twoZZ), twoXY - twoWZ, twoXZ + twoWY, twoXY + twoWZ, 1.0 - (twoXX + twoZZ),
		             twoYZ - twoWX, twoXZ - twoWY, twoYZ + twoWX, 1.0 - (twoXX + twoYY));
	}
	// [FCL-recalled C.1] Quaternion3f::fromRotation (trace / largest-diagonal branches)
	void fromRotation(const Matrix3f &R) {
		const int next[3] = {1, 2, 0};
		const FCL_REAL trace = R(0, 0) + R(1, 1) + R(2, 2);
		FCL_REAL root;
		if (trace > 0.0) {
			root = std::sqrt(trace + 1.0);
			q[0] = 0.5 * root;
			root = 0.5 / root;
			q[1] = (R(2, 1) - R(1, 2)) * root;
			q[2] = (R(0, 2) - R(2, 0)) * root;
			q[3] = (R(1, 0) - R(0, 1)) * root;
		} else {
			int i = 0;
			if (R(1, 1) > R(0, 0)) i = 1;
			if (R(2, 2) > R(i, i)) i = 2;
			const int j = next[i], k = next[j];
			root = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0);
			FCL_REAL *quat[3] = {&q[1], &q[2], &q[3]};
			*quat[i] = 0.5 * root;
			root = 0.5 / root;
			q[0] = (R(k, j) - R(j, k)) * root;
			*quat[j] = (R(j, i) + R(i, j)) * root;
			*quat[k] = (R(k, i) + R(i, k)) * root;
		}
	}

private:
	std::array<FCL_REAL, 4> q;
};

class Transform3f {
public:
	Transform3f() : matrix_set(true) {}
	Transform3f(const Matrix3f &R_, const Vec3f &T_) : matrix_set(true), R(R_), T(T_) { q.fromRotation(R_); }
	Transform3f(const Quaternion3f &q_, const Vec3f &T_) : matrix_set(false), T(T_), q(q_) { q.toRotation(R); }
	explicit Transform3f(const Matrix3f &R_) : matrix_set(true), R(R_) { q.fromRotation(R_); }
	explicit Transform3f(const Quaternion3f &q_) : matrix_set(false), q(q_) { q.toRotation(R); }
	explicit Transform3f(const Vec3f &T_) : matrix_set(true), T(T_) {}  // identity matrix AND identity quaternion

	const Vec3f &getTranslation() const { return T; }
	const Matrix3f &getRotation() const { return R; }
	const Quaternion3f &getQuatRotation() const { return q; }
	void setTransform(const Matrix3f &R_, const Vec3f &T_) {
		R = R_;
		T = T_;
		matrix_set = true;
		q.fromRotation(R_);
	}
	void setTransform(const Quaternion3f &q_, const Vec3f &T_) {
		matrix_set = false;
		q = q_;
		T = T_;
		q.toRotation(R);
	}
	void setRotation(const Matrix3f &R_) {
		R = R_;
		matrix_set = true;
		q.fromRotation(R_);
	}
	void setTranslation(const Vec3f &T_) { T = T_; }
	void setQuatRotation(const Quaternion3f &q_) {
		matrix_set = false;
		q = q_;
		q.toRotation(R);
	}
	void setIdentity() { *this = Transform3f(); }
	bool isIdentity() const { return q.isIdentity() && T[0] == 0 && T[1] == 0 && T[2] == 0; }
	// FCL 0.3: transform(v) = q.transform(v) + T
	Vec3f transform(const Vec3f &v) const { return q.transform(v) + T; }

	// which representation the caller supplied (decides the libmptcc pose format, see the header comment)
	bool builtFromQuaternion() const { return !matrix_set; }

private:
	bool matrix_set;
	Matrix3f R;
	Vec3f T;
	Quaternion3f q;
};

}  // namespace fcl
