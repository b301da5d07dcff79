/* cand_study.c -- research tool (NOT product, NOT oracle): per broadphase candidate (pose, object), the number of OBB
 * tests its depth-first traversal needs (nearest child first, until its first contact) next to features the
 * broadphase kernel could compute, to see whether long candidates can be told from short ones before they start.
 *   gcc -O2 -shared -fPIC -o /tmp/cstudy.so tools/research/cand_study.c -lm -lpthread */
#include "heuristic_study.c"

/* out: 6 doubles per candidate: pose, object, n_bv, hit, relative penetration of the two AABBs (min over the axes of
 * overlap length / agent box length), object triangles.  Returns the number of candidates (<= cap). */
uint64_t cand_study(const orc_env *env, const orc_model *link, uint64_t n, const double *poses, int fmt, uint64_t cap, double *out) {
	const int rl = rec_len(fmt);
	uint64_t nc = 0;
	for (uint64_t p = 0; p < n; ++p) {
		xform at;
		aabb ab;
		xform_decode(poses + (size_t)rl * p, fmt, &at);
		object_world_aabb(link, &at, &ab);
		for (int e = 0; e < env->n; ++e) {
			if (!aabb_overlap(&env->box[e], &ab)) continue;
			htrav tv;
			tv.m2 = link;
			tv.mode = 1;
			tv.hit = 0;
			tv.n_bv = tv.n_leaf = 0;
			tv.m1 = env->models[e];
			relative_transform(&env->tf[e], &at, &tv.R, tv.T);
			double sc;
			tv.n_bv++;
			if (obb_pair(&tv, tv.m1->bvs, tv.m2->bvs, &sc)) hrecurse(&tv, 0, 0);
			double pen = 1e300;
			for (int k = 0; k < 3; ++k) {
				const double lo = fmax(env->box[e].mn[k], ab.mn[k]), hi = fmin(env->box[e].mx[k], ab.mx[k]);
				const double r = (hi - lo) / fmax(ab.mx[k] - ab.mn[k], 1e-300);
				if (r < pen) pen = r;
			}
			if (nc < cap) {
				double *o = out + 6 * nc;
				o[0] = (double)p; o[1] = e; o[2] = (double)tv.n_bv; o[3] = tv.hit; o[4] = pen; o[5] = (double)env->models[e]->nt;
			}
			++nc;
		}
	}
	return nc;
}
