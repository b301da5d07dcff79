/* warp_sim.c -- research tool (NOT product, NOT oracle): a lock-step simulation of the narrowphase kernel's
 * SCHEDULING on the real geometry -- 32 lanes per warp, one depth-first plan stack per lane, two OBB tests per lane
 * step, a shared leaf queue tested 32 pairs at a time -- to count warp steps, OBB tests and lane utilisation for
 * different ways of handing candidates to lanes:
 *   mode 0 : batches of B candidates per warp; lanes steal pending plans once the batch has no unstarted candidate
 *            (the round-1/round-2 kernels)
 *   mode 1 : streaming -- an idle lane takes the next candidate of the global list; stealing only once the list is
 *            exhausted
 * Includes the oracle's source for its data structures and exact tests (fp64 tests stand in for the fp32 filter).
 *   gcc -O2 -shared -fPIC -o /tmp/wsim.so tools/research/warp_sim.c -lm -lpthread */
#include "../../oracle/fcl_oracle.c"

typedef struct {
	int pose, obj;
	mat3 R;
	double T[3];
} wcand;

typedef struct {
	int x, y;    /* env node, agent node */
	int kind;    /* 0 single (root pair), 1 children are env nodes x,x+1 vs y, 2 children are agent nodes y,y+1 vs x */
} wplan;

#define STK 128
typedef struct {
	wplan stk[STK];
	int top;
	wplan cur;
	int has;
	int cand; /* candidate of everything this lane holds */
	int paused;
	int age;    /* node steps on the current candidate */
	int last_s; /* stationary node of the lane's previous step: kind * 2^24 + id, -1 = none */
} wlane;

typedef struct {
	int envp, agp, cand;
} wleaf;

typedef struct {
	wlane lane[32];
	wleaf lq[512];
	int nlq;
	int qhead, qend; /* mode 0: unstarted candidates [qhead, qend) of the batch */
	int done;
	uint64_t steps_node, steps_leaf, active_lanes, stolen, same_s, lane_steps;
} wwarp;

static int wtest(const orc_model *m1, const orc_model *m2, const wcand *c, int b1, int b2, double *d2) {
	const bvnode *n1 = m1->bvs + b1, *n2 = m2->bvs + b2;
	mat3 R0b2, B;
	for (int r = 0; r < 3; ++r)
		for (int k = 0; k < 3; ++k) R0b2.m[r][k] = dot3(c->R.m[r], n2->axis[k]);
	for (int r = 0; r < 3; ++r)
		for (int k = 0; k < 3; ++k) B.m[r][k] = tdot(&R0b2, k, n1->axis[r]);
	double Rt[3], Tt[3], T[3];
	mat_vec(&c->R, n2->To, Rt);
	for (int k = 0; k < 3; ++k) Tt[k] = Rt[k] + c->T[k] - n1->To[k];
	for (int k = 0; k < 3; ++k) T[k] = dot3(Tt, n1->axis[k]);
	*d2 = T[0] * T[0] + T[1] * T[1] + T[2] * T[2];
	return !obb_disjoint(&B, T, n1->extent, n2->extent);
}

typedef struct {
	int mode, batch, leaf_thresh, n_warps, pause_on_leaf, steal_in_stream;
	int steal_age; /* > 0: an idle lane first takes the top pending plan of a lane that has spent >= steal_age steps on its
	                  candidate, and only otherwise the next candidate of the list */
	int cta_warps; /* > 1: once the list is exhausted, idle lanes also take pending plans from the other warps of their CTA */
	uint64_t rounds, steps_node, steps_leaf, tests, leaf_tests, active_lanes, stolen, same_s, lane_steps;
} wsim_cfg;

int wsim_run(const orc_env *env, const orc_model *link, uint64_t n, const double *poses, int fmt, wsim_cfg *cfg) {
	const int rl = rec_len(fmt);
	/* candidate list, pose-major */
	size_t cap = 1024, nc = 0;
	wcand *cl = malloc(cap * sizeof(wcand));
	for (uint64_t p = 0; p < n; ++p) {
		xform at;
		aabb ab;
		xform_decode(poses + (size_t)rl * p, fmt, &at);
		object_world_aabb(link, &at, &ab);
		for (int e = 0; e < env->n; ++e) {
			if (!aabb_overlap(&env->box[e], &ab)) continue;
			if (nc == cap) cl = realloc(cl, (cap *= 2) * sizeof(wcand));
			cl[nc].pose = (int)p;
			cl[nc].obj = e;
			relative_transform(&env->tf[e], &at, &cl[nc].R, cl[nc].T);
			++nc;
		}
	}
	unsigned char *dead = calloc(n, 1);
	wwarp *W = calloc(cfg->n_warps, sizeof(wwarp));
	size_t next = 0;
	uint64_t rounds = 0, tests = 0, leaf_tests = 0;
	int alive = cfg->n_warps;
	for (int w = 0; w < cfg->n_warps; ++w)
		for (int l = 0; l < 32; ++l) W[w].lane[l].cand = -1;
	while (alive > 0) {
		++rounds;
		for (int w = 0; w < cfg->n_warps; ++w) {
			wwarp *ww = &W[w];
			if (ww->done) continue;
			/* pop */
			for (int l = 0; l < 32; ++l) {
				wlane *L = &ww->lane[l];
				if (!L->has && L->top > 0 && !L->paused) {
					L->cur = L->stk[--L->top];
					L->has = 1;
				}
			}
			/* acquire */
			int n_idle = 0;
			for (int l = 0; l < 32; ++l) n_idle += !ww->lane[l].has && !ww->lane[l].paused;
			if (cfg->mode == 0) {
				if (n_idle == 32 && ww->nlq == 0 && ww->qhead >= ww->qend) {
					int anyp = 0;
					for (int l = 0; l < 32; ++l) anyp |= ww->lane[l].paused;
					if (!anyp) {
						/* batch finished: take the next one */
						if (next >= nc) {
							ww->done = 1;
							--alive;
							continue;
						}
						ww->qhead = (int)next;
						next += (size_t)cfg->batch;
						if (next > nc) next = nc;
						ww->qend = (int)next;
					}
				}
				for (int l = 0; l < 32; ++l) {
					wlane *L = &ww->lane[l];
					if (!L->has && !L->paused && ww->qhead < ww->qend) {
						L->cand = ww->qhead++;
						L->last_s = -1;
						L->cur = (wplan){0, 0, 0};
						L->has = 1;
						L->top = 0;
					}
				}
			} else {
				if (cfg->steal_age > 0) {
					for (int l = 0; l < 32; ++l) {
						wlane *L = &ww->lane[l];
						if (L->has || L->paused) continue;
						int best = -1;
						for (int d2 = 0; d2 < 32; ++d2) {
							wlane *D = &ww->lane[d2];
							if (D->has && D->top > 0 && D->age >= cfg->steal_age && (best < 0 || D->age > ww->lane[best].age)) best = d2;
						}
						if (best < 0) break;
						wlane *D = &ww->lane[best];
						L->cur = D->stk[--D->top];
						L->cand = D->cand;
						L->age = D->age;
						L->last_s = -1;
						L->has = 1;
						L->top = 0;
						ww->stolen++;
					}
				}
				for (int l = 0; l < 32; ++l) {
					wlane *L = &ww->lane[l];
					if (!L->has && !L->paused && next < nc) {
						L->cand = (int)next++;
						L->last_s = -1;
						L->age = 0;
						L->cur = (wplan){0, 0, 0};
						L->has = 1;
						L->top = 0;
					}
				}
			}
			const int may_steal = cfg->mode == 0 ? ww->qhead >= ww->qend : (next >= nc || cfg->steal_in_stream);
			if (may_steal) {
				int d = 0;
				for (int l = 0; l < 32; ++l) {
					wlane *L = &ww->lane[l];
					if (L->has || L->paused) continue;
					while (d < 32 && !(ww->lane[d].has && ww->lane[d].top > 0)) ++d;
					if (d == 32) break;
					L->cur = ww->lane[d].stk[--ww->lane[d].top];
					L->cand = ww->lane[d].cand;
					L->last_s = -1;
					L->has = 1;
					L->top = 0;
					ww->stolen++;
					++d;
				}
			}
			if (cfg->cta_warps > 1 && cfg->mode == 1 && next >= nc) {
				const int c0 = w / cfg->cta_warps * cfg->cta_warps;
				for (int l = 0; l < 32; ++l) {
					wlane *L = &ww->lane[l];
					if (L->has || L->paused) continue;
					int found = 0;
					for (int w2 = c0; w2 < c0 + cfg->cta_warps && w2 < cfg->n_warps && !found; ++w2) {
						if (w2 == w || W[w2].done) continue;
						for (int d2 = 0; d2 < 32; ++d2) {
							wlane *D = &W[w2].lane[d2];
							if (D->has && D->top > 0) {
								L->cur = D->stk[--D->top];
								L->cand = D->cand;
								L->last_s = -1;
								L->has = 1;
								L->top = 0;
								ww->stolen++;
								found = 1;
								break;
							}
						}
					}
				}
			}
			int n_act = 0;
			for (int l = 0; l < 32; ++l) n_act += ww->lane[l].has;
			if (n_act == 0 && ww->nlq == 0) {
				int anyp = 0;
				for (int l = 0; l < 32; ++l) anyp |= ww->lane[l].paused;
				if (cfg->mode == 1 && next >= nc && !anyp) {
					int others = 0;
					if (cfg->cta_warps > 1) {
						const int c0 = w / cfg->cta_warps * cfg->cta_warps;
						for (int w2 = c0; w2 < c0 + cfg->cta_warps && w2 < cfg->n_warps; ++w2) {
							if (w2 == w || W[w2].done) continue;
							for (int d2 = 0; d2 < 32; ++d2) others |= W[w2].lane[d2].has || W[w2].lane[d2].top > 0;
							others |= W[w2].nlq > 0;
						}
					}
					if (!others) {
						ww->done = 1;
						--alive;
					}
				}
				if (!anyp) continue;
			}
			if (ww->nlq >= cfg->leaf_thresh || (n_act == 0 && ww->nlq > 0)) {
				/* leaf phase */
				int k = ww->nlq < 32 ? ww->nlq : 32;
				ww->steps_leaf++;
				for (int i = 0; i < k; ++i) {
					const wleaf lf = ww->lq[--ww->nlq];
					const wcand *c = &cl[lf.cand];
					if (dead[c->pose]) continue;
					++leaf_tests;
					const orc_model *m1 = env->models[c->obj];
					const uint32_t *t1 = m1->tri + 3 * (size_t)lf.envp, *t2 = link->tri + 3 * (size_t)lf.agp;
					if (tri_intersect(m1->xyz + 3 * (size_t)t1[0], m1->xyz + 3 * (size_t)t1[1], m1->xyz + 3 * (size_t)t1[2],
					                  link->xyz + 3 * (size_t)t2[0], link->xyz + 3 * (size_t)t2[1], link->xyz + 3 * (size_t)t2[2],
					                  &c->R, c->T))
						dead[c->pose] = 1;
				}
				if (cfg->pause_on_leaf) {
					/* a lane stays paused while one of its pairs is still queued */
					for (int l = 0; l < 32; ++l) {
						wlane *L = &ww->lane[l];
						if (!L->paused) continue;
						int still = 0;
						for (int i = 0; i < ww->nlq; ++i) still |= ww->lq[i].cand == L->cand;
						if (!still) L->paused = 0;
					}
				}
				continue;
			}
			/* node step */
			ww->steps_node++;
			ww->active_lanes += (uint64_t)n_act;
			for (int l = 0; l < 32; ++l) {
				wlane *L = &ww->lane[l];
				if (!L->has) continue;
				L->has = 0;
				L->age++;
				const wplan pl = L->cur;
				const wcand *c = &cl[L->cand];
				if (dead[c->pose]) {
					L->top = 0;
					continue;
				}
				{
					const int sid = pl.kind == 2 ? (2 << 24) + pl.x : (1 << 24) + pl.y;
					ww->lane_steps++;
					if (sid == L->last_s) ww->same_s++;
					L->last_s = sid;
				}
				const orc_model *m1 = env->models[c->obj];
				wplan out[2];
				int ok[2] = {0, 0};
				double d2[2] = {0, 0};
				const int nk = pl.kind == 0 ? 1 : 2;
				for (int k = 0; k < nk; ++k) {
					const int i1 = pl.x + (pl.kind == 1 ? k : 0), i2 = pl.y + (pl.kind == 2 ? k : 0);
					++tests;
					if (!wtest(m1, link, c, i1, i2, &d2[k])) continue;
					const bvnode *n1 = m1->bvs + i1, *n2 = link->bvs + i2;
					const int l1 = n1->first_child < 0, l2 = n2->first_child < 0;
					if (l1 && l2) {
						ww->lq[ww->nlq++] = (wleaf){-(n1->first_child + 1), -(n2->first_child + 1), L->cand};
						if (cfg->pause_on_leaf) L->paused = 1;
						continue;
					}
					const int first = l2 || (!l1 && sqr_length3(n1->extent) > sqr_length3(n2->extent));
					out[k] = first ? (wplan){n1->first_child, i2, 1} : (wplan){i1, n2->first_child, 2};
					ok[k] = 1;
				}
				if (ok[0] && ok[1]) {
					const int sw = d2[1] < d2[0];
					L->cur = sw ? out[1] : out[0];
					L->has = 1;
					L->stk[L->top++] = sw ? out[0] : out[1];
				} else if (ok[0]) {
					L->cur = out[0];
					L->has = 1;
				} else if (ok[1]) {
					L->cur = out[1];
					L->has = 1;
				}
				if (L->paused && L->has) {  /* keep the plan but do not run it while paused */
					L->stk[L->top++] = L->cur;
					L->has = 0;
				}
			}
		}
	}
	cfg->rounds = rounds;
	cfg->tests = tests;
	cfg->leaf_tests = leaf_tests;
	cfg->steps_node = cfg->steps_leaf = cfg->active_lanes = cfg->stolen = cfg->same_s = cfg->lane_steps = 0;
	for (int w = 0; w < cfg->n_warps; ++w) {
		cfg->steps_node += W[w].steps_node;
		cfg->steps_leaf += W[w].steps_leaf;
		cfg->active_lanes += W[w].active_lanes;
		cfg->stolen += W[w].stolen;
		cfg->same_s += W[w].same_s;
		cfg->lane_steps += W[w].lane_steps;
	}
	free(W);
	free(dead);
	free(cl);
	return (int)nc;
}
