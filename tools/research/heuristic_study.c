/* heuristic_study.c -- research tool (NOT product, NOT oracle): how many OBB tests does a depth-first BVH-vs-BVH
 * traversal need until the first contact, as a function of the ORDER in which overlapping children are visited?
 * The verdict does not depend on that order, the work does.  Includes the oracle's source for its data structures
 * and exact tests.   gcc -O2 -shared -fPIC -o /tmp/heur.so tools/research/heuristic_study.c -lm -lpthread */
#include "../../oracle/fcl_oracle.c"

typedef struct {
	const orc_model *m1, *m2;
	mat3 R;
	double T[3];
	int mode, hit;
	uint64_t n_bv, n_leaf;
} htrav;

/* overlap test + a score (larger = visit first) */
static int obb_pair(const htrav *tv, const bvnode *b1, const bvnode *b2, double *score) {
	mat3 R0b2, B;
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c) R0b2.m[r][c] = dot3(tv->R.m[r], b2->axis[c]);
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c) B.m[r][c] = tdot(&R0b2, c, b1->axis[r]);
	double Rt[3], Tt[3], T[3];
	mat_vec(&tv->R, b2->To, Rt);
	for (int k = 0; k < 3; ++k) Tt[k] = Rt[k] + tv->T[k] - b1->To[k];
	for (int k = 0; k < 3; ++k) T[k] = dot3(Tt, b1->axis[k]);
	if (obb_disjoint(&B, T, b1->extent, b2->extent)) return 0;
	const double *a = b1->extent, *b = b2->extent;
	double d2 = T[0] * T[0] + T[1] * T[1] + T[2] * T[2];
	/* relative slack on the six face axes: (ra + rb - |t|) / (ra + rb) */
	double mins = 1e300, sum = 0;
	for (int k = 0; k < 3; ++k) {
		double rr = a[k] + fabs(B.m[k][0]) * b[0] + fabs(B.m[k][1]) * b[1] + fabs(B.m[k][2]) * b[2];
		double s = (rr - fabs(T[k])) / (rr + 1e-300);
		if (s < mins) mins = s;
		sum += s;
		double tb = B.m[0][k] * T[0] + B.m[1][k] * T[1] + B.m[2][k] * T[2];
		rr = b[k] + fabs(B.m[0][k]) * a[0] + fabs(B.m[1][k]) * a[1] + fabs(B.m[2][k]) * a[2];
		s = (rr - fabs(tb)) / (rr + 1e-300);
		if (s < mins) mins = s;
		sum += s;
	}
	double sz = sqr_length3(a) + sqr_length3(b);
	switch (tv->mode) {
	case 1: *score = -d2; break;                 /* nearest centres first (the kernel today) */
	case 2: *score = mins; break;                /* deepest overlap (min relative slack over the face axes) */
	case 3: *score = sum; break;                 /* largest mean relative slack */
	case 4: *score = -d2 / (sz + 1e-300); break; /* centre distance relative to the boxes' size */
	default: *score = 0; break;
	}
	return 1;
}

static void hrecurse(htrav *tv, int b1, int b2) {
	const bvnode *n1 = tv->m1->bvs + b1, *n2 = tv->m2->bvs + b2;
	const int l1 = n1->first_child < 0, l2 = n2->first_child < 0;
	if (l1 && l2) {
		tv->n_leaf++;
		const uint32_t *t1 = tv->m1->tri + 3 * (size_t)(-(n1->first_child + 1));
		const uint32_t *t2 = tv->m2->tri + 3 * (size_t)(-(n2->first_child + 1));
		const double *v1 = tv->m1->xyz, *v2 = tv->m2->xyz;
		if (tri_intersect(v1 + 3 * (size_t)t1[0], v1 + 3 * (size_t)t1[1], v1 + 3 * (size_t)t1[2], v2 + 3 * (size_t)t2[0],
		                  v2 + 3 * (size_t)t2[1], v2 + 3 * (size_t)t2[2], &tv->R, tv->T))
			tv->hit = 1;
		return;
	}
	const double sz1 = sqr_length3(n1->extent), sz2 = sqr_length3(n2->extent);
	const int first = l2 || (!l1 && (sz1 > sz2));
	int ca[2], cb[2], ok[2];
	double sc[2];
	for (int k = 0; k < 2; ++k) {
		ca[k] = first ? n1->first_child + k : b1;
		cb[k] = first ? b2 : n2->first_child + k;
		tv->n_bv++;
		ok[k] = obb_pair(tv, tv->m1->bvs + ca[k], tv->m2->bvs + cb[k], &sc[k]);
	}
	const int o0 = (ok[0] && ok[1] && tv->mode != 0 && sc[1] > sc[0]) ? 1 : 0;
	for (int t = 0; t < 2 && !tv->hit; ++t) {
		const int k = t == 0 ? o0 : 1 - o0;
		if (ok[k]) hrecurse(tv, ca[k], cb[k]);
	}
}

/* per pose-set: BV tests and leaf tests until the first contact (single-link agent); candidates in object order */
int heur_count(const orc_env *env, const orc_model *link, uint64_t n, const double *poses, int fmt, int mode,
               uint64_t *bv, uint64_t *leaf, unsigned char *hit) {
	const int rl = rec_len(fmt);
	for (uint64_t p = 0; p < n; ++p) {
		xform at;
		aabb ab;
		xform_decode(poses + (size_t)rl * p, fmt, &at);
		object_world_aabb(link, &at, &ab);
		htrav tv;
		tv.m2 = link;
		tv.mode = mode;
		tv.hit = 0;
		tv.n_bv = tv.n_leaf = 0;
		for (int e = 0; e < env->n && !tv.hit; ++e) {
			if (!aabb_overlap(&env->box[e], &ab)) continue;
			tv.m1 = env->models[e];
			relative_transform(&env->tf[e], &at, &tv.R, tv.T);
			double sc;
			tv.n_bv++;
			if (obb_pair(&tv, tv.m1->bvs, tv.m2->bvs, &sc)) hrecurse(&tv, 0, 0);
		}
		bv[p] = tv.n_bv;
		leaf[p] = tv.n_leaf;
		hit[p] = (unsigned char)tv.hit;
	}
	return 0;
}
