/*
 * fcl_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See fcl_oracle.h.
 *
 * PARITY UNPINNED (no FCL in this image, no golden vectors in the reference).  Every function
 * cites the reference call site it serves and the upstream FCL 0.3.x file whose published
 * algorithm it restates ([FCL-recalled]).  All arithmetic is IEEE
 * double, evaluated left-to-right as written; build with -ffp-contract=off so that no FMA is
 * formed (FCL on x86-64 without -march=native has none).
 */
#include "fcl_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------- */
/* C.1  Vec3f / Matrix3f / Quaternion3f / Transform3f  (fcl/math/vec_3f.h, matrix_3f.h,          */
/*      src/math/transform.cpp)                                                                  */
/* ------------------------------------------------------------------------------------------- */

typedef struct { double v[3]; } vec3;
typedef struct { double m[3][3]; } mat3; /* row-major, m[r][c] */

static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static inline void cross3(const double *l, const double *r, double *o) {
	double x = l[1] * r[2] - l[2] * r[1];
	double y = l[2] * r[0] - l[0] * r[2];
	double z = l[0] * r[1] - l[1] * r[0];
	o[0] = x; o[1] = y; o[2] = z;
}

static inline double sqr_length3(const double *a) { return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; }

/* Matrix3f::transposeDotX/Y/Z(v): column c of M dotted with v */
static inline double tdot(const mat3 *M, int c, const double *v) {
	return M->m[0][c] * v[0] + M->m[1][c] * v[1] + M->m[2][c] * v[2];
}

/* Matrix3f * Vec3f = (dotX(v), dotY(v), dotZ(v)) */
static inline void mat_vec(const mat3 *M, const double *v, double *o) {
	double x = dot3(M->m[0], v), y = dot3(M->m[1], v), z = dot3(M->m[2], v);
	o[0] = x; o[1] = y; o[2] = z;
}

typedef struct {
	mat3 R;      /* rotation matrix (Transform3f::R, materialised)  */
	double T[3]; /* translation                                     */
	double q[4]; /* w x y z                                         */
} xform;

/* Quaternion3f::toRotation */
static void quat_to_rotation(const double *q, mat3 *R) {
	double twoX = 2.0 * q[1], twoY = 2.0 * q[2], twoZ = 2.0 * q[3];
	double twoWX = twoX * q[0], twoWY = twoY * q[0], twoWZ = twoZ * q[0];
	double twoXX = twoX * q[1], twoXY = twoY * q[1], twoXZ = twoZ * q[1];
	double twoYY = twoY * q[2], twoYZ = twoZ * q[2], twoZZ = twoZ * q[3];
	R->m[0][0] = 1.0 - (twoYY + twoZZ); R->m[0][1] = twoXY - twoWZ;         R->m[0][2] = twoXZ + twoWY;
	R->m[1][0] = twoXY + twoWZ;         R->m[1][1] = 1.0 - (twoXX + twoZZ); R->m[1][2] = twoYZ - twoWX;
	R->m[2][0] = twoXZ - twoWY;         R->m[2][1] = twoYZ + twoWX;         R->m[2][2] = 1.0 - (twoXX + twoYY);
}

/* Quaternion3f::fromRotation (Ogre-style branch on the trace) */
static void quat_from_rotation(const mat3 *R, double *q) {
	static const int next[3] = {1, 2, 0};
	double trace = R->m[0][0] + R->m[1][1] + R->m[2][2];
	double root;
	if (trace > 0.0) {
		root = sqrt(trace + 1.0);
		q[0] = 0.5 * root;
		root = 0.5 / root;
		q[1] = (R->m[2][1] - R->m[1][2]) * root;
		q[2] = (R->m[0][2] - R->m[2][0]) * root;
		q[3] = (R->m[1][0] - R->m[0][1]) * root;
	} else {
		int i = 0;
		if (R->m[1][1] > R->m[0][0]) i = 1;
		if (R->m[2][2] > R->m[i][i]) i = 2;
		int j = next[i], k = next[j];
		root = sqrt(R->m[i][i] - R->m[j][j] - R->m[k][k] + 1.0);
		q[1 + i] = 0.5 * root;
		root = 0.5 / root;
		q[0] = (R->m[k][j] - R->m[j][k]) * root;
		q[1 + j] = (R->m[j][i] + R->m[i][j]) * root;
		q[1 + k] = (R->m[k][i] + R->m[i][k]) * root;
	}
}

/* Quaternion3f::operator* (Hamilton product, term order as upstream) */
static void quat_mul(const double *a, const double *b, double *o) {
	double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
	double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
	double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
	double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
	o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}

/* Quaternion3f::transform(v) = (q * (0,v) * conj(q)).xyz */
static void quat_transform(const double *q, const double *v, double *o) {
	double p[4] = {0.0, v[0], v[1], v[2]};
	double c[4] = {q[0], -q[1], -q[2], -q[3]};
	double t[4], r[4];
	quat_mul(q, p, t);
	quat_mul(t, c, r);
	o[0] = r[1]; o[1] = r[2]; o[2] = r[3];
}

static inline int quat_is_identity(const double *q) { return q[0] == 1.0 && q[1] == 0.0 && q[2] == 0.0 && q[3] == 0.0; }

/* Pose record -> Transform3f.
 *  RT12: Transform3f(R,T) keeps R,T and sets q.fromRotation(R); Transform3f(T) (identity R, as
 *        Omnidirectional::getPoses builds, /root/reference/agents/omnidirectional.hpp:224-242) gives the
 *        same q = (1,0,0,0) through fromRotation(I), so one record format covers both constructors.
 *  TQ7 : Transform3f(Quaternion3f(w,x,y,z), Vec3f(x,y,z)) keeps q and derives R = q.toRotation()
 *        (fcl_helpers::parseTransform, /root/reference/utilities/fcl_helpers.hpp:16-25; PRMLite samples,
 *        /root/reference/discretizations/workspace/prmlite.hpp:246-258). */
static void xform_decode(const double *rec, int fmt, xform *x) {
	if (fmt == ORC_POSE_TQ7) {
		x->T[0] = rec[0]; x->T[1] = rec[1]; x->T[2] = rec[2];
		x->q[0] = rec[3]; x->q[1] = rec[4]; x->q[2] = rec[5]; x->q[3] = rec[6];
		quat_to_rotation(x->q, &x->R);
	} else {
		for (int r = 0; r < 3; ++r)
			for (int c = 0; c < 3; ++c) x->R.m[r][c] = rec[3 * r + c];
		x->T[0] = rec[9]; x->T[1] = rec[10]; x->T[2] = rec[11];
		quat_from_rotation(&x->R, x->q);
	}
}

void orc_pose_decode(const double *rec, int fmt, double *R, double *T, double *q) {
	xform x;
	xform_decode(rec, fmt, &x);
	if (R) memcpy(R, x.R.m, sizeof(double) * 9);
	if (T) memcpy(T, x.T, sizeof(double) * 3);
	if (q) memcpy(q, x.q, sizeof(double) * 4);
}

/* ------------------------------------------------------------------------------------------- */
/* C.2-C.4  BVHModel<OBBRSS> build  (src/BVH/BVH_model.cpp, BV_fitter.cpp, BV_splitter.cpp,     */
/*          BVH_utility.cpp)                                                                     */
/* ------------------------------------------------------------------------------------------- */

typedef struct {
	double axis[3][3]; /* obb.axis[k] */
	double To[3];      /* obb.To      */
	double extent[3];  /* obb.extent  */
	int first_child;   /* >=0 inner (children first_child, first_child+1); <0 leaf: prim = -(fc+1) */
	int first_primitive;
	int num_primitives;
} bvnode;

struct orc_model {
	uint32_t nv, nt;
	double *xyz;      /* [nv][3] */
	uint32_t *tri;    /* [nt][3] */
	bvnode *bvs;      /* [2nt-1] */
	uint32_t *prim;   /* primitive_indices */
	int num_bvs;
	/* computeLocalAABB */
	double aabb_min[3], aabb_max[3], aabb_center[3], aabb_radius;
};

/* getCovariance (BVH_utility.cpp): covariance of the 3n triangle vertices */
static void get_covariance(const orc_model *m, const uint32_t *idx, int n, mat3 *M) {
	double S1[3] = {0, 0, 0};
	double S2[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
	for (int i = 0; i < n; ++i) {
		const uint32_t *t = m->tri + 3 * (size_t)idx[i];
		const double *p1 = m->xyz + 3 * (size_t)t[0];
		const double *p2 = m->xyz + 3 * (size_t)t[1];
		const double *p3 = m->xyz + 3 * (size_t)t[2];
		S1[0] += (p1[0] + p2[0] + p3[0]);
		S1[1] += (p1[1] + p2[1] + p3[1]);
		S1[2] += (p1[2] + p2[2] + p3[2]);
		S2[0][0] += (p1[0] * p1[0] + p2[0] * p2[0] + p3[0] * p3[0]);
		S2[1][1] += (p1[1] * p1[1] + p2[1] * p2[1] + p3[1] * p3[1]);
		S2[2][2] += (p1[2] * p1[2] + p2[2] * p2[2] + p3[2] * p3[2]);
		S2[0][1] += (p1[0] * p1[1] + p2[0] * p2[1] + p3[0] * p3[1]);
		S2[0][2] += (p1[0] * p1[2] + p2[0] * p2[2] + p3[0] * p3[2]);
		S2[1][2] += (p1[1] * p1[2] + p2[1] * p2[2] + p3[1] * p3[2]);
	}
	int n_points = 3 * n;
	M->m[0][0] = S2[0][0] - S1[0] * S1[0] / n_points;
	M->m[1][1] = S2[1][1] - S1[1] * S1[1] / n_points;
	M->m[2][2] = S2[2][2] - S1[2] * S1[2] / n_points;
	M->m[0][1] = S2[0][1] - S1[0] * S1[1] / n_points;
	M->m[1][2] = S2[1][2] - S1[1] * S1[2] / n_points;
	M->m[0][2] = S2[0][2] - S1[0] * S1[2] / n_points;
	M->m[1][0] = M->m[0][1];
	M->m[2][0] = M->m[0][2];
	M->m[2][1] = M->m[1][2];
}

/* eigen() of matrix_3f.h: Numerical-Recipes cyclic Jacobi, <= 50 sweeps.  vout[k] = ROW k of v. */
static void eigen_jacobi(const mat3 *min, double dout[3], double vout[3][3]) {
	mat3 R = *min;
	const int n = 3;
	int j, iq, ip, i;
	double tresh, theta, tau, t, sm, s, h, g, c;
	double b[3], z[3], d[3];
	double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};

	for (ip = 0; ip < n; ++ip) {
		b[ip] = d[ip] = R.m[ip][ip];
		z[ip] = 0;
	}
	for (i = 0; i < 50; ++i) {
		sm = 0;
		for (ip = 0; ip < n; ++ip)
			for (iq = ip + 1; iq < n; ++iq) sm += fabs(R.m[ip][iq]);
		if (sm == 0.0) {
			for (int k = 0; k < 3; ++k) {
				vout[k][0] = v[k][0]; vout[k][1] = v[k][1]; vout[k][2] = v[k][2];
				dout[k] = d[k];
			}
			return;
		}
		if (i < 3) tresh = 0.2 * sm / (n * n);
		else tresh = 0.0;
		for (ip = 0; ip < n; ++ip) {
			for (iq = ip + 1; iq < n; ++iq) {
				g = 100.0 * fabs(R.m[ip][iq]);
				if (i > 3 && fabs(d[ip]) + g == fabs(d[ip]) && fabs(d[iq]) + g == fabs(d[iq])) {
					R.m[ip][iq] = 0.0;
				} else if (fabs(R.m[ip][iq]) > tresh) {
					h = d[iq] - d[ip];
					if (fabs(h) + g == fabs(h)) {
						t = (R.m[ip][iq]) / h;
					} else {
						theta = 0.5 * h / (R.m[ip][iq]);
						t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
						if (theta < 0.0) t = -t;
					}
					c = 1.0 / sqrt(1 + t * t);
					s = t * c;
					tau = s / (1.0 + c);
					h = t * R.m[ip][iq];
					z[ip] -= h;
					z[iq] += h;
					d[ip] -= h;
					d[iq] += h;
					R.m[ip][iq] = 0.0;
					for (j = 0; j < ip; ++j) {
						g = R.m[j][ip]; h = R.m[j][iq];
						R.m[j][ip] = g - s * (h + g * tau);
						R.m[j][iq] = h + s * (g - h * tau);
					}
					for (j = ip + 1; j < iq; ++j) {
						g = R.m[ip][j]; h = R.m[j][iq];
						R.m[ip][j] = g - s * (h + g * tau);
						R.m[j][iq] = h + s * (g - h * tau);
					}
					for (j = iq + 1; j < n; ++j) {
						g = R.m[ip][j]; h = R.m[iq][j];
						R.m[ip][j] = g - s * (h + g * tau);
						R.m[iq][j] = h + s * (g - h * tau);
					}
					for (j = 0; j < n; ++j) {
						g = v[j][ip]; h = v[j][iq];
						v[j][ip] = g - s * (h + g * tau);
						v[j][iq] = h + s * (g - h * tau);
					}
				}
			}
		}
		for (ip = 0; ip < n; ++ip) {
			b[ip] += z[ip];
			d[ip] = b[ip];
			z[ip] = 0.0;
		}
	}
	/* "eigen: too many iterations in Jacobi transform": upstream returns with vout default-constructed
	 * (zero vectors) and dout untouched; zero both here. */
	for (int k = 0; k < 3; ++k) {
		vout[k][0] = vout[k][1] = vout[k][2] = 0.0;
		dout[k] = 0.0;
	}
}

/* BVFitter<OBBRSS>::fit (OBB half only; the RSS half is never read by collision queries) */
static void fit_obb(const orc_model *m, const uint32_t *idx, int n, bvnode *bv) {
	mat3 M;
	double E[3][3], s[3];
	get_covariance(m, idx, n, &M);
	eigen_jacobi(&M, s, E);

	int min, mid, max;
	if (s[0] > s[1]) { max = 0; min = 1; }
	else { min = 0; max = 1; }
	if (s[2] < s[min]) { mid = min; min = 2; }
	else if (s[2] > s[max]) { mid = max; max = 2; }
	else { mid = 2; }
	(void)min;

	bv->axis[0][0] = E[0][max]; bv->axis[0][1] = E[1][max]; bv->axis[0][2] = E[2][max];
	bv->axis[1][0] = E[0][mid]; bv->axis[1][1] = E[1][mid]; bv->axis[1][2] = E[2][mid];
	bv->axis[2][0] = E[1][max] * E[2][mid] - E[1][mid] * E[2][max];
	bv->axis[2][1] = E[0][mid] * E[2][max] - E[0][max] * E[2][mid];
	bv->axis[2][2] = E[0][max] * E[1][mid] - E[0][mid] * E[1][max];

	/* getExtentAndCenter_mesh */
	double min_coord[3] = {DBL_MAX, DBL_MAX, DBL_MAX};
	double max_coord[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
	for (int i = 0; i < n; ++i) {
		const uint32_t *t = m->tri + 3 * (size_t)idx[i];
		for (int j = 0; j < 3; ++j) {
			const double *p = m->xyz + 3 * (size_t)t[j];
			double proj[3];
			proj[0] = dot3(bv->axis[0], p);
			proj[1] = dot3(bv->axis[1], p);
			proj[2] = dot3(bv->axis[2], p);
			for (int k = 0; k < 3; ++k) {
				if (proj[k] > max_coord[k]) max_coord[k] = proj[k];
				if (proj[k] < min_coord[k]) min_coord[k] = proj[k];
			}
		}
	}
	double o[3] = {(max_coord[0] + min_coord[0]) / 2, (max_coord[1] + min_coord[1]) / 2,
	               (max_coord[2] + min_coord[2]) / 2};
	for (int c = 0; c < 3; ++c) bv->To[c] = bv->axis[0][c] * o[0] + bv->axis[1][c] * o[1] + bv->axis[2][c] * o[2];
	bv->extent[0] = (max_coord[0] - min_coord[0]) / 2;
	bv->extent[1] = (max_coord[1] - min_coord[1]) / 2;
	bv->extent[2] = (max_coord[2] - min_coord[2]) / 2;
}

/* BVHModel::recursiveBuildTree with BVSplitter<OBBRSS> SPLIT_METHOD_MEAN */
static void recursive_build(orc_model *m, int bv_id, int first_primitive, int num_primitives) {
	bvnode *node = m->bvs + bv_id;
	uint32_t *cur = m->prim + first_primitive;

	fit_obb(m, cur, num_primitives, node);

	/* computeRule_mean: split_vector = obb.axis[0]; split_value = (sum of vertex sums . sv)/(3n) */
	double sv[3] = {node->axis[0][0], node->axis[0][1], node->axis[0][2]};
	double c[3] = {0.0, 0.0, 0.0};
	for (int i = 0; i < num_primitives; ++i) {
		const uint32_t *t = m->tri + 3 * (size_t)cur[i];
		const double *p1 = m->xyz + 3 * (size_t)t[0];
		const double *p2 = m->xyz + 3 * (size_t)t[1];
		const double *p3 = m->xyz + 3 * (size_t)t[2];
		c[0] += (p1[0] + p2[0] + p3[0]);
		c[1] += (p1[1] + p2[1] + p3[1]);
		c[2] += (p1[2] + p2[2] + p3[2]);
	}
	double split_value = (c[0] * sv[0] + c[1] * sv[1] + c[2] * sv[2]) / (3 * num_primitives);

	node->first_primitive = first_primitive;
	node->num_primitives = num_primitives;

	if (num_primitives == 1) {
		node->first_child = -((int)cur[0] + 1);
		return;
	}
	node->first_child = m->num_bvs;
	m->num_bvs += 2;

	int c1 = 0;
	for (int i = 0; i < num_primitives; ++i) {
		const uint32_t *t = m->tri + 3 * (size_t)cur[i];
		const double *p1 = m->xyz + 3 * (size_t)t[0];
		const double *p2 = m->xyz + 3 * (size_t)t[1];
		const double *p3 = m->xyz + 3 * (size_t)t[2];
		double p[3];
		p[0] = (p1[0] + p2[0] + p3[0]) / 3.0;
		p[1] = (p1[1] + p2[1] + p3[1]) / 3.0;
		p[2] = (p1[2] + p2[2] + p3[2]) / 3.0;
		if (dot3(sv, p) > split_value) {
			/* right side: stays */
		} else {
			uint32_t tmp = cur[i];
			cur[i] = cur[c1];
			cur[c1] = tmp;
			c1++;
		}
	}
	if (c1 == 0 || c1 == num_primitives) c1 = num_primitives / 2;

	int fc = node->first_child; /* node pointer stays valid: bvs is preallocated */
	recursive_build(m, fc, first_primitive, c1);
	recursive_build(m, fc + 1, first_primitive + c1, num_primitives - c1);
}

/* BVHModel::computeLocalAABB */
static void compute_local_aabb(orc_model *m) {
	double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
	for (uint32_t i = 0; i < m->nv; ++i) {
		const double *p = m->xyz + 3 * (size_t)i;
		for (int k = 0; k < 3; ++k) {
			if (p[k] < mn[k]) mn[k] = p[k];
			if (p[k] > mx[k]) mx[k] = p[k];
		}
	}
	double ctr[3] = {(mn[0] + mx[0]) * 0.5, (mn[1] + mx[1]) * 0.5, (mn[2] + mx[2]) * 0.5};
	double rad = 0;
	for (uint32_t i = 0; i < m->nv; ++i) {
		const double *p = m->xyz + 3 * (size_t)i;
		double d[3] = {ctr[0] - p[0], ctr[1] - p[1], ctr[2] - p[2]};
		double r = sqr_length3(d);
		if (r > rad) rad = r;
	}
	memcpy(m->aabb_min, mn, sizeof mn);
	memcpy(m->aabb_max, mx, sizeof mx);
	memcpy(m->aabb_center, ctr, sizeof ctr);
	m->aabb_radius = sqrt(rad);
}

/* beginModel / addSubModel(verts, tris) / endModel, as called from
 * /root/reference/utilities/meshhandler.hpp:44-47 and :130-133 */
orc_model *orc_model_build(const double *xyz, uint32_t nv, const uint32_t *tri, uint32_t nt) {
	if (nv == 0 || nt == 0) return NULL;
	orc_model *m = (orc_model *)calloc(1, sizeof *m);
	m->nv = nv;
	m->nt = nt;
	m->xyz = (double *)malloc(sizeof(double) * 3 * (size_t)nv);
	m->tri = (uint32_t *)malloc(sizeof(uint32_t) * 3 * (size_t)nt);
	memcpy(m->xyz, xyz, sizeof(double) * 3 * (size_t)nv);
	memcpy(m->tri, tri, sizeof(uint32_t) * 3 * (size_t)nt);
	m->bvs = (bvnode *)calloc(2 * (size_t)nt - 1, sizeof(bvnode));
	m->prim = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)nt);
	for (uint32_t i = 0; i < nt; ++i) m->prim[i] = i;
	m->num_bvs = 1;
	recursive_build(m, 0, 0, (int)nt);
	compute_local_aabb(m);
	return m;
}

void orc_model_free(orc_model *m) {
	if (!m) return;
	free(m->xyz); free(m->tri); free(m->bvs); free(m->prim); free(m);
}

uint32_t orc_model_num_nodes(const orc_model *m) { return (uint32_t)m->num_bvs; }
uint32_t orc_model_num_tris(const orc_model *m) { return m->nt; }

void orc_model_get_nodes(const orc_model *m, double *axis, double *To, double *extent, int32_t *first_child,
                         int32_t *first_primitive, int32_t *num_primitives) {
	for (int i = 0; i < m->num_bvs; ++i) {
		const bvnode *b = m->bvs + i;
		if (axis) memcpy(axis + 9 * (size_t)i, b->axis, sizeof(double) * 9);
		if (To) memcpy(To + 3 * (size_t)i, b->To, sizeof(double) * 3);
		if (extent) memcpy(extent + 3 * (size_t)i, b->extent, sizeof(double) * 3);
		if (first_child) first_child[i] = b->first_child;
		if (first_primitive) first_primitive[i] = b->first_primitive;
		if (num_primitives) num_primitives[i] = b->num_primitives;
	}
}

void orc_model_get_primitive_indices(const orc_model *m, uint32_t *out) { memcpy(out, m->prim, sizeof(uint32_t) * m->nt); }

void orc_model_get_local_aabb(const orc_model *m, double *aabb6, double *center3, double *radius) {
	if (aabb6) { memcpy(aabb6, m->aabb_min, 24); memcpy(aabb6 + 3, m->aabb_max, 24); }
	if (center3) memcpy(center3, m->aabb_center, 24);
	if (radius) *radius = m->aabb_radius;
}

/* ------------------------------------------------------------------------------------------- */
/* C.5  CollisionObject(cgeom, tf): computeLocalAABB + computeAABB  (fcl/collision_object.h)     */
/* ------------------------------------------------------------------------------------------- */

typedef struct { double mn[3], mx[3]; } aabb;

static void object_world_aabb(const orc_model *m, const xform *t, aabb *out) {
	if (quat_is_identity(t->q)) {
		for (int k = 0; k < 3; ++k) {
			out->mn[k] = m->aabb_min[k] + t->T[k];
			out->mx[k] = m->aabb_max[k] + t->T[k];
		}
	} else {
		/* Transform3f::transform(v) = q.transform(v) + T */
		double c[3];
		quat_transform(t->q, m->aabb_center, c);
		for (int k = 0; k < 3; ++k) {
			double ck = c[k] + t->T[k];
			out->mn[k] = ck - m->aabb_radius;
			out->mx[k] = ck + m->aabb_radius;
		}
	}
}

void orc_object_aabb(const orc_model *m, const double *rec, int fmt, double *aabb6) {
	xform x;
	aabb a;
	xform_decode(rec, fmt, &x);
	object_world_aabb(m, &x, &a);
	memcpy(aabb6, a.mn, 24);
	memcpy(aabb6 + 3, a.mx, 24);
}

static inline int aabb_overlap(const aabb *a, const aabb *b) {
	if (a->mn[0] > b->mx[0]) return 0;
	if (a->mn[1] > b->mx[1]) return 0;
	if (a->mn[2] > b->mx[2]) return 0;
	if (a->mx[0] < b->mn[0]) return 0;
	if (a->mx[1] < b->mn[1]) return 0;
	if (a->mx[2] < b->mn[2]) return 0;
	return 1;
}

static inline double aabb_size(const aabb *a) {
	double d[3] = {a->mx[0] - a->mn[0], a->mx[1] - a->mn[1], a->mx[2] - a->mn[2]};
	return sqr_length3(d);
}

static inline void aabb_merge(const aabb *a, const aabb *b, aabb *o) {
	for (int k = 0; k < 3; ++k) {
		o->mn[k] = a->mn[k] < b->mn[k] ? a->mn[k] : b->mn[k];
		o->mx[k] = a->mx[k] > b->mx[k] ? a->mx[k] : b->mx[k];
	}
}

/* ------------------------------------------------------------------------------------------- */
/* C.6  DynamicAABBTreeCollisionManager: registerObjects -> HierarchyTree::init_0 -> topdown_0   */
/*      (src/broadphase/broadphase_dynamic_AABB_tree.cpp, fcl/broadphase/hierarchy_tree.hxx).    */
/*      registerObjects() on an empty manager bulk-builds the tree and marks it set up, so the   */
/*      following setup() is a no-op.  Tree shape only fixes the ORDER of candidate pairs.       */
/* ------------------------------------------------------------------------------------------- */

typedef struct dnode {
	aabb bv;
	struct dnode *child[2]; /* child[1] == NULL => leaf */
	int obj;                /* leaf: object index        */
} dnode;

typedef struct {
	dnode *pool;
	int n_pool, n_used;
	dnode *root;
	int n_leaves;
} dtree;

static inline double aabb_center_k(const aabb *a, int k) { return (a->mn[k] + a->mx[k]) * 0.5; }

/* std::nth_element substitute (quickselect on bv.center()[axis]); any valid nth_element partition
 * is a faithful FCL tree, and the verdict does not depend on it. */
static void select_nth(dnode **a, int n, int nth, int axis) {
	int lo = 0, hi = n - 1;
	while (lo < hi) {
		double pivot = aabb_center_k(&a[(lo + hi) / 2]->bv, axis);
		int i = lo, j = hi;
		while (i <= j) {
			while (aabb_center_k(&a[i]->bv, axis) < pivot) ++i;
			while (pivot < aabb_center_k(&a[j]->bv, axis)) --j;
			if (i <= j) {
				dnode *t = a[i]; a[i] = a[j]; a[j] = t;
				++i; --j;
			}
		}
		if (nth <= j) hi = j;
		else if (nth >= i) lo = i;
		else return;
	}
}

static dnode *dtree_new_node(dtree *t) { return &t->pool[t->n_used++]; }

static dnode *dtree_topdown(dtree *t, dnode **lbeg, int num_leaves) {
	const int bu_threshold = 2;
	if (num_leaves > 1) {
		if (num_leaves > bu_threshold) {
			aabb vol = lbeg[0]->bv;
			for (int i = 1; i < num_leaves; ++i) aabb_merge(&vol, &lbeg[i]->bv, &vol);
			int best_axis = 0;
			double extent[3] = {vol.mx[0] - vol.mn[0], vol.mx[1] - vol.mn[1], vol.mx[2] - vol.mn[2]};
			if (extent[1] > extent[0]) best_axis = 1;
			if (extent[2] > extent[best_axis]) best_axis = 2;
			int half = num_leaves / 2;
			select_nth(lbeg, num_leaves, half, best_axis);
			dnode *node = dtree_new_node(t);
			node->bv = vol;
			node->obj = -1;
			node->child[0] = dtree_topdown(t, lbeg, half);
			node->child[1] = dtree_topdown(t, lbeg + half, num_leaves - half);
			return node;
		} else {
			/* bottomup() with two leaves: one parent over both */
			dnode *node = dtree_new_node(t);
			aabb_merge(&lbeg[0]->bv, &lbeg[1]->bv, &node->bv);
			node->obj = -1;
			node->child[0] = lbeg[0];
			node->child[1] = lbeg[1];
			return node;
		}
	}
	return lbeg[0];
}

static void dtree_init(dtree *t, const aabb *boxes, int n) {
	t->n_leaves = n;
	t->n_used = 0;
	t->root = NULL;
	if (n == 0) { t->pool = NULL; t->n_pool = 0; return; }
	t->n_pool = 2 * n;
	t->pool = (dnode *)malloc(sizeof(dnode) * (size_t)t->n_pool);
	dnode **leaves = (dnode **)malloc(sizeof(dnode *) * (size_t)n);
	for (int i = 0; i < n; ++i) {
		dnode *l = dtree_new_node(t);
		l->bv = boxes[i];
		l->child[0] = l->child[1] = NULL;
		l->obj = i;
		leaves[i] = l;
	}
	t->root = dtree_topdown(t, leaves, n);
	free(leaves);
}

static void dtree_free(dtree *t) { free(t->pool); t->pool = NULL; }

/* ------------------------------------------------------------------------------------------- */
/* C.8  OBB overlap (src/BV/OBB.cpp: overlap(R0,T0,b1,b2) -> obbDisjoint(B,T,a,b))              */
/* ------------------------------------------------------------------------------------------- */

static int obb_disjoint(const mat3 *B, const double *T, const double *a, const double *b) {
	double t, s;
	const double reps = 1e-6;
	mat3 Bf;
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c) Bf.m[r][c] = fabs(B->m[r][c]) + reps;

	/* A0 */
	t = ((T[0] < 0.0) ? -T[0] : T[0]);
	if (t > (a[0] + dot3(Bf.m[0], b))) return 1;
	/* B0 */
	s = tdot(B, 0, T);
	t = ((s < 0.0) ? -s : s);
	if (t > (b[0] + tdot(&Bf, 0, a))) return 1;
	/* A1 */
	t = ((T[1] < 0.0) ? -T[1] : T[1]);
	if (t > (a[1] + dot3(Bf.m[1], b))) return 1;
	/* A2 */
	t = ((T[2] < 0.0) ? -T[2] : T[2]);
	if (t > (a[2] + dot3(Bf.m[2], b))) return 1;
	/* B1 */
	s = tdot(B, 1, T);
	t = ((s < 0.0) ? -s : s);
	if (t > (b[1] + tdot(&Bf, 1, a))) return 1;
	/* B2 */
	s = tdot(B, 2, T);
	t = ((s < 0.0) ? -s : s);
	if (t > (b[2] + tdot(&Bf, 2, a))) return 1;

	/* A0 x B0 */
	s = T[2] * B->m[1][0] - T[1] * B->m[2][0];
	t = ((s < 0.0) ? -s : s);
	if (t > (a[1] * Bf.m[2][0] + a[2] * Bf.m[1][0] + b[1] * Bf.m[0][2] + b[2] * Bf.m[0][1])) return 1;
	/* A0 x B1 */
	s = T[2] * B->m[1][1] - T[1] * B->m[2][1];
	t = ((s < 0.0) ? -s : s);
	if (t > (a[1] * Bf.m[2][1] + a[2] * Bf.m[1][1] + b[0] * Bf.m[0][2] + b[2] * Bf.m[0][0])) return 1;
	/* A0 x B2 */
	s = T[2] * B->m[1][2] - T[1] * B->m[2][2];
	t = ((s < 0.0) ? -s : s);
	if (t > (a[1] * Bf.m[2][2] + a[2] * Bf.m[1][2] + b[0] * Bf.m[0][1] + b[1] * Bf.m[0][0])) return 1;
	/* A1 x B0 */
	s = T[0] * B->m[2][0] - T[2] * B->m[0][0];
	t = ((s < 0.0) ? -s : s);
	if (t > (a[0] * Bf.m[2][0] + a[2] * Bf.m[0][0] + b[1] * Bf.m[1][2] + b[2] * Bf.m[1][1])) return 1;
	/* A1 x B1 */
	s = T[0] * B->m[2][1] - T[2] * B->m[0][1];
	t = ((s < 0.0) ? -s : s);
	if (t > (a[0] * Bf.m[2][1] + a[2] * Bf.m[0][1] + b[0] * Bf.m[1][2] + b[2] * Bf.m[1][0])) return 1;
	/* A1 x B2 */
	s = T[0] * B->m[2][2] - T[2] * B->m[0][2];
	t = ((s < 0.0) ? -s : s);
	if (t > (a[0] * Bf.m[2][2] + a[2] * Bf.m[0][2] + b[0] * Bf.m[1][1] + b[1] * Bf.m[1][0])) return 1;
	/* A2 x B0 */
	s = T[1] * B->m[0][0] - T[0] * B->m[1][0];
	t = ((s < 0.0) ? -s : s);
	if (t > (a[0] * Bf.m[1][0] + a[1] * Bf.m[0][0] + b[1] * Bf.m[2][2] + b[2] * Bf.m[2][1])) return 1;
	/* A2 x B1 */
	s = T[1] * B->m[0][1] - T[0] * B->m[1][1];
	t = ((s < 0.0) ? -s : s);
	if (t > (a[0] * Bf.m[1][1] + a[1] * Bf.m[0][1] + b[0] * Bf.m[2][2] + b[2] * Bf.m[2][0])) return 1;
	/* A2 x B2 */
	s = T[1] * B->m[0][2] - T[0] * B->m[1][2];
	t = ((s < 0.0) ? -s : s);
	if (t > (a[0] * Bf.m[1][2] + a[1] * Bf.m[0][2] + b[0] * Bf.m[2][1] + b[1] * Bf.m[2][0])) return 1;

	return 0;
}

static int obb_overlap(const mat3 *R0, const double *T0, const bvnode *b1, const bvnode *b2) {
	mat3 R0b2, B;
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c) R0b2.m[r][c] = dot3(R0->m[r], b2->axis[c]);
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c) B.m[r][c] = tdot(&R0b2, c, b1->axis[r]);
	double Rt[3], Ttemp[3], T[3];
	mat_vec(R0, b2->To, Rt);
	for (int k = 0; k < 3; ++k) Ttemp[k] = Rt[k] + T0[k] - b1->To[k];
	T[0] = dot3(Ttemp, b1->axis[0]);
	T[1] = dot3(Ttemp, b1->axis[1]);
	T[2] = dot3(Ttemp, b1->axis[2]);
	return !obb_disjoint(&B, T, b1->extent, b2->extent);
}

int orc_obb_overlap(const double *R0, const double *T0, const double *axis1, const double *To1, const double *ext1,
                    const double *axis2, const double *To2, const double *ext2) {
	bvnode a, b;
	mat3 R;
	memcpy(R.m, R0, 72);
	memcpy(a.axis, axis1, 72); memcpy(a.To, To1, 24); memcpy(a.extent, ext1, 24);
	memcpy(b.axis, axis2, 72); memcpy(b.To, To2, 24); memcpy(b.extent, ext2, 24);
	return obb_overlap(&R, T0, &a, &b);
}

/* ------------------------------------------------------------------------------------------- */
/* C.9  Intersect::intersect_Triangle + project6  (src/intersect.cpp)                            */
/* ------------------------------------------------------------------------------------------- */

/* std::min / std::max semantics */
static inline double dmin(double a, double b) { return (b < a) ? b : a; }
static inline double dmax(double a, double b) { return (a < b) ? b : a; }

static inline int project6(const double *ax, const double *p1, const double *p2, const double *p3,
                           const double *q1, const double *q2, const double *q3) {
	double P1 = dot3(ax, p1), P2 = dot3(ax, p2), P3 = dot3(ax, p3);
	double Q1 = dot3(ax, q1), Q2 = dot3(ax, q2), Q3 = dot3(ax, q3);
	double mn1 = dmin(P1, dmin(P2, P3));
	double mx2 = dmax(Q1, dmax(Q2, Q3));
	if (mn1 > mx2) return 0;
	double mx1 = dmax(P1, dmax(P2, P3));
	double mn2 = dmin(Q1, dmin(Q2, Q3));
	if (mn2 > mx1) return 0;
	return 1;
}

static inline void sub3(const double *a, const double *b, double *o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }

static int tri_intersect(const double *P1, const double *P2, const double *P3, const double *Q1in,
                         const double *Q2in, const double *Q3in, const mat3 *R, const double *T) {
	double Q1[3], Q2[3], Q3[3];
	mat_vec(R, Q1in, Q1); mat_vec(R, Q2in, Q2); mat_vec(R, Q3in, Q3);
	for (int k = 0; k < 3; ++k) { Q1[k] = Q1[k] + T[k]; Q2[k] = Q2[k] + T[k]; Q3[k] = Q3[k] + T[k]; }

	double p1[3], p2[3], p3[3], q1[3], q2[3], q3[3];
	sub3(P1, P1, p1); sub3(P2, P1, p2); sub3(P3, P1, p3);
	sub3(Q1, P1, q1); sub3(Q2, P1, q2); sub3(Q3, P1, q3);

	double e1[3], e2[3], e3[3], f1[3], f2[3], f3[3], n1[3], m1[3], ax[3];
	sub3(p2, p1, e1);
	sub3(p3, p2, e2);
	cross3(e1, e2, n1);
	if (!project6(n1, p1, p2, p3, q1, q2, q3)) return 0;

	sub3(q2, q1, f1);
	sub3(q3, q2, f2);
	cross3(f1, f2, m1);
	if (!project6(m1, p1, p2, p3, q1, q2, q3)) return 0;

	cross3(e1, f1, ax);
	if (!project6(ax, p1, p2, p3, q1, q2, q3)) return 0;
	cross3(e1, f2, ax);
	if (!project6(ax, p1, p2, p3, q1, q2, q3)) return 0;
	sub3(q1, q3, f3);
	cross3(e1, f3, ax);
	if (!project6(ax, p1, p2, p3, q1, q2, q3)) return 0;

	cross3(e2, f1, ax);
	if (!project6(ax, p1, p2, p3, q1, q2, q3)) return 0;
	cross3(e2, f2, ax);
	if (!project6(ax, p1, p2, p3, q1, q2, q3)) return 0;
	cross3(e2, f3, ax);
	if (!project6(ax, p1, p2, p3, q1, q2, q3)) return 0;

	sub3(p1, p3, e3);
	cross3(e3, f1, ax);
	if (!project6(ax, p1, p2, p3, q1, q2, q3)) return 0;
	cross3(e3, f2, ax);
	if (!project6(ax, p1, p2, p3, q1, q2, q3)) return 0;
	cross3(e3, f3, ax);
	if (!project6(ax, p1, p2, p3, q1, q2, q3)) return 0;

	cross3(e1, n1, ax);
	if (!project6(ax, p1, p2, p3, q1, q2, q3)) return 0;
	cross3(e2, n1, ax);
	if (!project6(ax, p1, p2, p3, q1, q2, q3)) return 0;
	cross3(e3, n1, ax);
	if (!project6(ax, p1, p2, p3, q1, q2, q3)) return 0;

	cross3(f1, m1, ax);
	if (!project6(ax, p1, p2, p3, q1, q2, q3)) return 0;
	cross3(f2, m1, ax);
	if (!project6(ax, p1, p2, p3, q1, q2, q3)) return 0;
	cross3(f3, m1, ax);
	if (!project6(ax, p1, p2, p3, q1, q2, q3)) return 0;

	return 1;
}

int orc_tri_intersect(const double *P, const double *Q, const double *R, const double *T) {
	mat3 Rm;
	memcpy(Rm.m, R, 72);
	return tri_intersect(P, P + 3, P + 6, Q, Q + 3, Q + 6, &Rm, T);
}

/* ------------------------------------------------------------------------------------------- */
/* C.7  fcl::collide -> orientedMeshCollide<MeshCollisionTraversalNodeOBBRSS> -> collisionRecurse */
/*      (src/collision.cpp, collision_func_matrix.cpp, traversal/traversal_recurse.cpp,           */
/*       traversal/traversal_node_setup.cpp, traversal/traversal_node_bvhs.cpp)                   */
/* ------------------------------------------------------------------------------------------- */

typedef struct {
	const orc_model *m1, *m2;
	mat3 R;
	double T[3];
	int num_contacts; /* CollisionResult::numContacts(), num_max_contacts = 1, enable_contact = false */
	orc_counters *ctr;
	/* test probe (orc_path_reaches): the traversal does not test triangles or stop, it only notes whether it reaches the
	 * leaf pair of these two primitives */
	int probe, probe_prim1, probe_prim2, probe_hit;
} trav;

static void collision_recurse(trav *tv, int b1, int b2) {
	const bvnode *n1 = tv->m1->bvs + b1;
	const bvnode *n2 = tv->m2->bvs + b2;
	int l1 = n1->first_child < 0, l2 = n2->first_child < 0;

	if (l1 && l2) {
		tv->ctr->n_bv_tests++;
		if (!obb_overlap(&tv->R, tv->T, n1, n2)) return;
		if (tv->probe) {
			if (-(n1->first_child + 1) == tv->probe_prim1 && -(n2->first_child + 1) == tv->probe_prim2) tv->probe_hit = 1;
			return;
		}
		/* meshCollisionOrientedNodeLeafTesting, enable_contact == false */
		tv->ctr->n_leaf_tests++;
		const uint32_t *t1 = tv->m1->tri + 3 * (size_t)(-(n1->first_child + 1));
		const uint32_t *t2 = tv->m2->tri + 3 * (size_t)(-(n2->first_child + 1));
		const double *v1 = tv->m1->xyz, *v2 = tv->m2->xyz;
		if (tri_intersect(v1 + 3 * (size_t)t1[0], v1 + 3 * (size_t)t1[1], v1 + 3 * (size_t)t1[2],
		                  v2 + 3 * (size_t)t2[0], v2 + 3 * (size_t)t2[1], v2 + 3 * (size_t)t2[2], &tv->R, tv->T)) {
			if (tv->num_contacts < 1) tv->num_contacts++;
		}
		return;
	}

	tv->ctr->n_bv_tests++;
	if (!obb_overlap(&tv->R, tv->T, n1, n2)) return;

	/* firstOverSecond: l2 || (!l1 && size1 > size2), size = obb.extent.sqrLength() */
	double sz1 = sqr_length3(n1->extent), sz2 = sqr_length3(n2->extent);
	if (l2 || (!l1 && (sz1 > sz2))) {
		int c1 = n1->first_child, c2 = n1->first_child + 1;
		collision_recurse(tv, c1, b2);
		if (tv->num_contacts >= 1) return; /* canStop() */
		collision_recurse(tv, c2, b2);
	} else {
		int c1 = n2->first_child, c2 = n2->first_child + 1;
		collision_recurse(tv, b1, c1);
		if (tv->num_contacts >= 1) return;
		collision_recurse(tv, b1, c2);
	}
}

/* relativeTransform(R1,T1,R2,T2): R = R1^T R2 (transposeTimes), T = R1^T (T2 - T1) */
static void relative_transform(const xform *t1, const xform *t2, mat3 *R, double *T) {
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c) {
			double col[3] = {t2->R.m[0][c], t2->R.m[1][c], t2->R.m[2][c]};
			R->m[r][c] = tdot(&t1->R, r, col);
		}
	double temp[3];
	sub3(t2->T, t1->T, temp);
	T[0] = tdot(&t1->R, 0, temp);
	T[1] = tdot(&t1->R, 1, temp);
	T[2] = tdot(&t1->R, 2, temp);
}

/* Does collisionRecurse on (m1, m2) under x1 = R x2 + T reach the leaf pair holding triangles (prim1, prim2), i.e. do all
 * the OBB tests on FCL's way down to it report an overlap?  (tests: the device code's path replay against it) */
int orc_path_reaches(const orc_model *m1, const orc_model *m2, const double *R, const double *T, int prim1, int prim2) {
	trav tv;
	orc_counters c;
	memset(&c, 0, sizeof c);
	tv.m1 = m1;
	tv.m2 = m2;
	memcpy(tv.R.m, R, sizeof(double) * 9);
	memcpy(tv.T, T, sizeof(double) * 3);
	tv.num_contacts = 0;
	tv.ctr = &c;
	tv.probe = 1;
	tv.probe_prim1 = prim1;
	tv.probe_prim2 = prim2;
	tv.probe_hit = 0;
	collision_recurse(&tv, 0, 0);
	return tv.probe_hit;
}

/* relativeTransform(tf1, tf2) of two pose records (tests: the device code's rel_element against it) */
void orc_relative_transform(const double *rec1, int fmt1, const double *rec2, int fmt2, double *R, double *T) {
	xform a, b;
	mat3 m;
	xform_decode(rec1, fmt1, &a);
	xform_decode(rec2, fmt2, &b);
	relative_transform(&a, &b, &m, T);
	memcpy(R, m.m, sizeof(double) * 9);
}

/* ------------------------------------------------------------------------------------------- */
/* StaticEnvironmentMeshHandler  (/root/reference/utilities/meshhandler.hpp:18-55)               */
/* ------------------------------------------------------------------------------------------- */

struct orc_env {
	int n;
	orc_model **models;
	xform *tf;
	aabb *box;
	dtree tree;
};

orc_env *orc_env_create(int n_objects, orc_model *const *models, const double *pose7) {
	orc_env *e = (orc_env *)calloc(1, sizeof *e);
	e->n = n_objects;
	e->models = (orc_model **)malloc(sizeof(orc_model *) * (size_t)(n_objects > 0 ? n_objects : 1));
	e->tf = (xform *)malloc(sizeof(xform) * (size_t)(n_objects > 0 ? n_objects : 1));
	e->box = (aabb *)malloc(sizeof(aabb) * (size_t)(n_objects > 0 ? n_objects : 1));
	for (int i = 0; i < n_objects; ++i) {
		e->models[i] = models[i];
		xform_decode(pose7 + 7 * (size_t)i, ORC_POSE_TQ7, &e->tf[i]);
		object_world_aabb(models[i], &e->tf[i], &e->box[i]); /* new fcl::CollisionObject(model, tf) */
	}
	dtree_init(&e->tree, e->box, n_objects); /* registerObjects(worldObjects); setup() */
	return e;
}

void orc_env_free(orc_env *e) {
	if (!e) return;
	dtree_free(&e->tree);
	free(e->models); free(e->tf); free(e->box); free(e);
}

void orc_env_get_object_aabb(const orc_env *e, int obj, double *aabb6) {
	memcpy(aabb6, e->box[obj].mn, 24);
	memcpy(aabb6 + 3, e->box[obj].mx, 24);
}

/* ------------------------------------------------------------------------------------------- */
/* MeshHandler::isInCollision  (/root/reference/utilities/meshhandler.hpp:187-243) and            */
/* fcl_helpers::defaultCollisionFunction (/root/reference/utilities/fcl_helpers.hpp:52-65)        */
/* ------------------------------------------------------------------------------------------- */

typedef struct {
	const orc_env *env;
	orc_model *const *links; /* model of agent object k = links[k % n_links] */
	int n_links;
	const xform *agent_tf;
	int num_contacts;
	int done;
	orc_counters *ctr;
} cdata;

/* defaultCollisionFunction(o1 = env object, o2 = agent object) */
static int collision_callback(cdata *cd, int env_obj, int agent_obj) {
	if (cd->done) return 1;
	/* fcl::collide(o1, o2, request, result): skipped inside orientedMeshCollide when the request is
	 * already satisfied */
	if (!(cd->num_contacts >= 1)) {
		trav tv;
		tv.m1 = cd->env->models[env_obj];
		tv.m2 = cd->links[agent_obj % cd->n_links];
		tv.num_contacts = cd->num_contacts;
		tv.probe = 0;
		tv.ctr = cd->ctr;
		cd->ctr->n_obj_pairs++;
		relative_transform(&cd->env->tf[env_obj], &cd->agent_tf[agent_obj], &tv.R, tv.T);
		collision_recurse(&tv, 0, 0);
		cd->num_contacts = tv.num_contacts;
	}
	if (cd->num_contacts >= 1) cd->done = 1;
	return cd->done;
}

/* collisionRecurse(root1, root2, cdata, callback) of broadphase_dynamic_AABB_tree.cpp */
static int broadphase_recurse(const dnode *r1, const dnode *r2, cdata *cd) {
	int leaf1 = r1->child[1] == NULL, leaf2 = r2->child[1] == NULL;
	if (leaf1 && leaf2) {
		cd->ctr->n_aabb_tests++;
		if (!aabb_overlap(&r1->bv, &r2->bv)) return 0;
		return collision_callback(cd, r1->obj, r2->obj);
	}
	cd->ctr->n_aabb_tests++;
	if (!aabb_overlap(&r1->bv, &r2->bv)) return 0;
	if (leaf2 || (!leaf1 && (aabb_size(&r1->bv) > aabb_size(&r2->bv)))) {
		if (broadphase_recurse(r1->child[0], r2, cd)) return 1;
		if (broadphase_recurse(r1->child[1], r2, cd)) return 1;
	} else {
		if (broadphase_recurse(r1, r2->child[0], cd)) return 1;
		if (broadphase_recurse(r1, r2->child[1], cd)) return 1;
	}
	return 0;
}

static int rec_len(int fmt) { return fmt == ORC_POSE_TQ7 ? 7 : 12; }

int orc_is_in_collision(const orc_env *env, int n_links, orc_model *const *links, uint64_t n_pose_sets,
                        const double *poses, int fmt, orc_counters *counters) {
	orc_counters local;
	memset(&local, 0, sizeof local);
	orc_counters *ctr = counters ? counters : &local;
	ctr->n_pose_sets++;

	size_t n_obj = (size_t)n_pose_sets * (size_t)n_links;
	if (n_obj == 0 || env->n == 0) return 0; /* collide() returns at once if either manager is empty */

	/* agent[j]->getMeshPose(pose[j]) == new fcl::CollisionObject(agentModel, tf): the constructor runs
	 * computeLocalAABB() over every model vertex each time (kept, so that the timing is faithful). */
	xform *tf = (xform *)malloc(sizeof(xform) * n_obj);
	aabb *box = (aabb *)malloc(sizeof(aabb) * n_obj);
	const int rl = rec_len(fmt);
	for (size_t k = 0; k < n_obj; ++k) {
		orc_model *m = links[k % (size_t)n_links];
		xform_decode(poses + (size_t)rl * k, fmt, &tf[k]);
		compute_local_aabb(m);
		object_world_aabb(m, &tf[k], &box[k]);
	}
	/* agentCollisionModel = new DynamicAABBTreeCollisionManager; registerObjects; setup */
	dtree at;
	dtree_init(&at, box, (int)n_obj);

	cdata cd;
	cd.env = env; cd.links = links; cd.n_links = n_links; cd.agent_tf = tf;
	cd.num_contacts = 0; cd.done = 0; cd.ctr = ctr;
	/* worldCollisionModel->collide(agentCollisionModel, &collisionData, defaultCollisionFunction) */
	broadphase_recurse(env->tree.root, at.root, &cd);

	dtree_free(&at);
	free(tf);
	free(box);
	return cd.num_contacts > 0;
}

typedef struct {
	const orc_env *env;
	int n_links;
	orc_model *const *links; /* per-thread private copies (compute_local_aabb writes the model) */
	const double *poses;
	int fmt;
	uint64_t begin, end;
	uint32_t *bits;
	int check_self;
	orc_counters ctr;
} slice_job;

/* checkSelfCollision (/root/reference/utilities/meshhandler.hpp:205-219): the links of ONE pose-set are put
 * in their own DynamicAABBTreeCollisionManager and collide(cdata, callback) runs selfCollisionRecurse: every
 * pair of leaves whose AABBs overlap (closed intervals) reaches defaultCollisionFunction until a contact is
 * found.  The verdict is "some overlapping pair of link meshes collides".
 * [deviation] Which link of a pair becomes object 1 depends, in FCL, on the tree that
 * HierarchyTree::topdown_0 builds with std::nth_element, whose partition is unspecified; this restatement
 * fixes the orientation to (lower link index, higher link index). */
int orc_self_collision(int n_links, orc_model *const *links, const double *pose_set, int fmt, orc_counters *counters) {
	orc_counters local;
	memset(&local, 0, sizeof local);
	orc_counters *ctr = counters ? counters : &local;
	if (n_links < 2) return 0;
	xform *tf = (xform *)malloc(sizeof(xform) * (size_t)n_links);
	aabb *box = (aabb *)malloc(sizeof(aabb) * (size_t)n_links);
	const int rl = fmt == ORC_POSE_TQ7 ? 7 : 12;
	for (int k = 0; k < n_links; ++k) {
		xform_decode(pose_set + (size_t)rl * (size_t)k, fmt, &tf[k]);
		compute_local_aabb(links[k]);
		object_world_aabb(links[k], &tf[k], &box[k]);
	}
	int hit = 0;
	for (int a = 0; a < n_links && !hit; ++a)
		for (int b = a + 1; b < n_links && !hit; ++b) {
			ctr->n_aabb_tests++;
			if (!aabb_overlap(&box[a], &box[b])) continue;
			trav tv;
			tv.m1 = links[a];
			tv.m2 = links[b];
			tv.num_contacts = 0;
			tv.probe = 0;
			tv.ctr = ctr;
			ctr->n_obj_pairs++;
			relative_transform(&tf[a], &tf[b], &tv.R, tv.T);
			collision_recurse(&tv, 0, 0);
			hit = tv.num_contacts > 0;
		}
	free(tf);
	free(box);
	return hit;
}

static void *slice_run(void *arg) {
	slice_job *j = (slice_job *)arg;
	const size_t stride = (size_t)rec_len(j->fmt) * (size_t)j->n_links;
	for (uint64_t p = j->begin; p < j->end; ++p) {
		/* meshhandler.hpp:205-219: the self-collision of a pose-set is checked first and settles the call */
		int hit = j->check_self && orc_self_collision(j->n_links, j->links, j->poses + stride * p, j->fmt, &j->ctr);
		if (!hit) hit = orc_is_in_collision(j->env, j->n_links, j->links, 1, j->poses + stride * p, j->fmt, &j->ctr);
		if (hit) j->bits[p >> 5] |= (1u << (p & 31));
	}
	return NULL;
}

/* Shallow per-thread copy of a model so that the faithful per-pose computeLocalAABB() does not race. */
static orc_model *model_shallow_copy(const orc_model *m) {
	orc_model *c = (orc_model *)malloc(sizeof *c);
	*c = *m;
	return c;
}

static int verdicts_impl(const orc_env *env, int n_links, orc_model *const *links, uint64_t n_pose_sets,
                         const double *poses, int fmt, uint32_t *verdict_bits, int n_threads, orc_counters *counters,
                         int check_self);

int orc_verdicts(const orc_env *env, int n_links, orc_model *const *links, uint64_t n_pose_sets,
                 const double *poses, int fmt, uint32_t *verdict_bits, int n_threads, orc_counters *counters) {
	return verdicts_impl(env, n_links, links, n_pose_sets, poses, fmt, verdict_bits, n_threads, counters, 0);
}

int orc_verdicts_self(const orc_env *env, int n_links, orc_model *const *links, uint64_t n_pose_sets,
                      const double *poses, int fmt, uint32_t *verdict_bits, int n_threads, orc_counters *counters) {
	return verdicts_impl(env, n_links, links, n_pose_sets, poses, fmt, verdict_bits, n_threads, counters, 1);
}

static int verdicts_impl(const orc_env *env, int n_links, orc_model *const *links, uint64_t n_pose_sets,
                         const double *poses, int fmt, uint32_t *verdict_bits, int n_threads, orc_counters *counters,
                         int check_self) {
	size_t n_words = (size_t)((n_pose_sets + 31) / 32);
	memset(verdict_bits, 0, sizeof(uint32_t) * n_words);
	if (n_threads < 1) n_threads = 1;
	if ((uint64_t)n_threads > n_words) n_threads = n_words ? (int)n_words : 1;

	slice_job *jobs = (slice_job *)calloc((size_t)n_threads, sizeof *jobs);
	pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof *th);
	size_t words_per = (n_words + (size_t)n_threads - 1) / (size_t)n_threads;
	for (int t = 0; t < n_threads; ++t) {
		orc_model **priv = (orc_model **)malloc(sizeof(orc_model *) * (size_t)n_links);
		for (int l = 0; l < n_links; ++l) priv[l] = model_shallow_copy(links[l]);
		uint64_t b = (uint64_t)t * words_per * 32, e = (uint64_t)(t + 1) * words_per * 32;
		if (b > n_pose_sets) b = n_pose_sets;
		if (e > n_pose_sets) e = n_pose_sets;
		jobs[t].env = env; jobs[t].n_links = n_links; jobs[t].links = priv; jobs[t].poses = poses;
		jobs[t].fmt = fmt; jobs[t].begin = b; jobs[t].end = e; jobs[t].bits = verdict_bits;
		jobs[t].check_self = check_self;
	}
	if (n_threads == 1) {
		slice_run(&jobs[0]);
	} else {
		for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, slice_run, &jobs[t]);
		for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
	}
	for (int t = 0; t < n_threads; ++t) {
		if (counters) {
			counters->n_pose_sets += jobs[t].ctr.n_pose_sets;
			counters->n_aabb_tests += jobs[t].ctr.n_aabb_tests;
			counters->n_obj_pairs += jobs[t].ctr.n_obj_pairs;
			counters->n_bv_tests += jobs[t].ctr.n_bv_tests;
			counters->n_leaf_tests += jobs[t].ctr.n_leaf_tests;
		}
		for (int l = 0; l < n_links; ++l) free((void *)jobs[t].links[l]);
		free((void *)jobs[t].links);
	}
	free(jobs);
	free(th);
	return 0;
}

int orc_verdicts_bruteforce(const orc_env *env, int n_links, orc_model *const *links, uint64_t n_pose_sets,
                            const double *poses, int fmt, uint32_t *verdict_bits) {
	size_t n_words = (size_t)((n_pose_sets + 31) / 32);
	memset(verdict_bits, 0, sizeof(uint32_t) * n_words);
	const int rl = rec_len(fmt);
	for (uint64_t p = 0; p < n_pose_sets; ++p) {
		int hit = 0;
		for (int l = 0; l < n_links && !hit; ++l) {
			const orc_model *am = links[l];
			xform at;
			aabb ab;
			xform_decode(poses + (size_t)rl * ((size_t)p * (size_t)n_links + (size_t)l), fmt, &at);
			object_world_aabb(am, &at, &ab);
			for (int e = 0; e < env->n && !hit; ++e) {
				if (!aabb_overlap(&env->box[e], &ab)) continue;
				const orc_model *em = env->models[e];
				mat3 R;
				double T[3];
				relative_transform(&env->tf[e], &at, &R, T);
				for (uint32_t i = 0; i < em->nt && !hit; ++i) {
					const uint32_t *t1 = em->tri + 3 * (size_t)i;
					for (uint32_t k = 0; k < am->nt; ++k) {
						const uint32_t *t2 = am->tri + 3 * (size_t)k;
						if (tri_intersect(em->xyz + 3 * (size_t)t1[0], em->xyz + 3 * (size_t)t1[1],
						                  em->xyz + 3 * (size_t)t1[2], am->xyz + 3 * (size_t)t2[0],
						                  am->xyz + 3 * (size_t)t2[1], am->xyz + 3 * (size_t)t2[2], &R, T)) {
							hit = 1;
							break;
						}
					}
				}
			}
		}
		if (hit) verdict_bits[p >> 5] |= (1u << (p & 31));
	}
	return 0;
}

/* ---- structure-free oracle, fast form ----------------------------------------------------------------------
 * Same verdict definition as orc_verdicts_bruteforce (exists object, link with overlapping world AABBs and a
 * triangle pair for which Intersect::intersect_Triangle holds), made fast enough for the deep trees (blimp merged
 * vs divider, the C5 scene) by two things that have nothing to do with the BVH, the OBB fit or the OBB test:
 *   - the pose range is split over pthreads;
 *   - a triangle pair is only handed to intersect_Triangle if the axis-aligned boxes of the two triangles (agent
 *     triangle moved into the object's frame with the same R, T) overlap after growing them by `pad` (1e-9 of the
 *     coordinates' magnitude).  Triangles whose boxes are further apart than that are separated along a coordinate
 *     axis by far more than fp64 round-off; if FCL's 17-axis test nevertheless reported a contact for such a pair,
 *     this oracle would say "free" where the BVH oracle says "collides" and the equality test would FAIL, i.e. the
 *     shortcut cannot hide a disagreement. */
typedef struct {
	const orc_env *env;
	int n_links;
	orc_model *const *links;
	const double *poses;
	int fmt;
	uint32_t *bits;
	uint64_t p0, p1;
	double **tri_box; /* per environment object: [nt][6] triangle boxes in the object's model frame */
} bf_slice;

static void *bf_run(void *arg) {
	bf_slice *s = (bf_slice *)arg;
	const int rl = rec_len(s->fmt);
	for (uint64_t p = s->p0; p < s->p1; ++p) {
		int hit = 0;
		for (int l = 0; l < s->n_links && !hit; ++l) {
			const orc_model *am = s->links[l];
			xform at;
			aabb ab;
			xform_decode(s->poses + (size_t)rl * ((size_t)p * (size_t)s->n_links + (size_t)l), s->fmt, &at);
			object_world_aabb(am, &at, &ab);
			for (int e = 0; e < s->env->n && !hit; ++e) {
				if (!aabb_overlap(&s->env->box[e], &ab)) continue;
				const orc_model *em = s->env->models[e];
				const double *tb = s->tri_box[e];
				mat3 R;
				double T[3];
				relative_transform(&s->env->tf[e], &at, &R, T);
				for (uint32_t k = 0; k < am->nt && !hit; ++k) {
					const uint32_t *t2 = am->tri + 3 * (size_t)k;
					double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX}, mag = 0.0;
					for (int c = 0; c < 3; ++c) {
						double q[3];
						mat_vec(&R, am->xyz + 3 * (size_t)t2[c], q);
						for (int a = 0; a < 3; ++a) {
							const double v = q[a] + T[a];
							if (v < lo[a]) lo[a] = v;
							if (v > hi[a]) hi[a] = v;
							if (fabs(v) > mag) mag = fabs(v);
						}
					}
					const double pad = 1e-9 * (mag + 1.0);
					for (uint32_t i = 0; i < em->nt; ++i) {
						const double *b = tb + 6 * (size_t)i;
						if (b[0] > hi[0] + pad || b[1] > hi[1] + pad || b[2] > hi[2] + pad || b[3] < lo[0] - pad ||
						    b[4] < lo[1] - pad || b[5] < lo[2] - pad)
							continue;
						const uint32_t *t1 = em->tri + 3 * (size_t)i;
						if (tri_intersect(em->xyz + 3 * (size_t)t1[0], em->xyz + 3 * (size_t)t1[1], em->xyz + 3 * (size_t)t1[2],
						                  am->xyz + 3 * (size_t)t2[0], am->xyz + 3 * (size_t)t2[1], am->xyz + 3 * (size_t)t2[2], &R,
						                  T)) {
							hit = 1;
							break;
						}
					}
				}
			}
		}
		if (hit) s->bits[p >> 5] |= (1u << (p & 31));
	}
	return NULL;
}

int orc_verdicts_bruteforce_fast(const orc_env *env, int n_links, orc_model *const *links, uint64_t n_pose_sets,
                                 const double *poses, int fmt, uint32_t *verdict_bits, int n_threads) {
	const size_t n_words = (size_t)((n_pose_sets + 31) / 32);
	memset(verdict_bits, 0, sizeof(uint32_t) * n_words);
	double **tri_box = (double **)calloc((size_t)(env->n > 0 ? env->n : 1), sizeof(double *));
	for (int e = 0; e < env->n; ++e) {
		const orc_model *em = env->models[e];
		int shared = -1; /* objects may share a model */
		for (int f = 0; f < e; ++f)
			if (env->models[f] == em) {
				shared = f;
				break;
			}
		if (shared >= 0) {
			tri_box[e] = tri_box[shared];
			continue;
		}
		tri_box[e] = (double *)malloc(sizeof(double) * 6 * (size_t)em->nt);
		for (uint32_t i = 0; i < em->nt; ++i) {
			double *b = tri_box[e] + 6 * (size_t)i;
			for (int a = 0; a < 3; ++a) {
				b[a] = DBL_MAX;
				b[3 + a] = -DBL_MAX;
			}
			for (int c = 0; c < 3; ++c) {
				const double *v = em->xyz + 3 * (size_t)em->tri[3 * (size_t)i + c];
				for (int a = 0; a < 3; ++a) {
					if (v[a] < b[a]) b[a] = v[a];
					if (v[a] > b[3 + a]) b[3 + a] = v[a];
				}
			}
		}
	}
	if (n_threads < 1) n_threads = 1;
	if (n_threads > 64) n_threads = 64;
	pthread_t th[64];
	int created[64];
	bf_slice sl[64];
	const uint64_t words_per = (n_words + (uint64_t)n_threads - 1) / (uint64_t)n_threads;
	for (int t = 0; t < n_threads; ++t) {
		uint64_t p0 = (uint64_t)t * words_per * 32, p1 = ((uint64_t)t + 1) * words_per * 32;
		if (p0 > n_pose_sets) p0 = n_pose_sets;
		if (p1 > n_pose_sets) p1 = n_pose_sets;
		sl[t] = (bf_slice){env, n_links, links, poses, fmt, verdict_bits, p0, p1, tri_box};
		created[t] = 0;
		if (p0 == p1) continue;
		if (t < n_threads - 1 && pthread_create(&th[t], NULL, bf_run, &sl[t]) == 0)
			created[t] = 1;
		else
			bf_run(&sl[t]); /* the last slice (or a failed create) runs on the calling thread */
	}
	for (int t = 0; t < n_threads; ++t)
		if (created[t]) pthread_join(th[t], NULL);
	for (int e = 0; e < env->n; ++e) {
		int shared = 0;
		for (int f = 0; f < e; ++f)
			if (env->models[f] == env->models[e]) shared = 1;
		if (!shared) free(tri_box[e]);
	}
	free(tri_box);
	return 0;
}
