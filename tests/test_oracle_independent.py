"""Checks of the oracle against things that do NOT share its derivation ("the oracle pins itself"):

  * the structure-free brute force (all triangle pairs, no BVH / OBB) on the DEEP trees -- blimp merged (1 355
    triangles, 21 levels) vs divider.dae, and a grazing-biased sample of the C5 scene found by bisection, which is
    where the argument "FCL's OBB test with its 1e-6 pad never prunes a pair its triangle test would accept" matters;
  * Intersect::intersect_Triangle against segment-triangle intersection by orientation predicates (extended precision);
  * obbDisjoint (Gottschalk's B-matrix algebra) against plain corner projection of the two boxes on the 15 axes.
"""
import numpy as np
import pytest

from helpers import Scene, env_bbox
from motionplanningtoolkit_b200 import meshio, scenes
from oracle import oracle as O


def test_bruteforce_equals_bvh_oracle_blimp_merge_vs_divider(built):
    env = meshio.environment_submeshes(meshio.load_mptm("divider"))
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    sc = Scene(env, [agent])
    rng = scenes.SplitMix64(0xB0B)
    lo, hi = env_bbox(env, 6.0)
    poses = np.concatenate([scenes.random_poses_near_objects(rng, 4000, env, 4.0),
                            scenes.random_poses_rt12(rng, 8000, lo, hi)])
    want = sc.oracle_bits(poses)
    frac = O.unpack_bits(want, len(poses)).mean()
    assert 0.2 < frac < 0.8
    assert np.array_equal(want, sc.brute_bits_fast(poses))
    tq = scenes.random_poses_tq7(rng, 3000, lo, hi)
    assert np.array_equal(sc.oracle_bits(tq, O.POSE_TQ7), sc.brute_bits_fast(tq, O.POSE_TQ7))
    # the fast form is the plain all-pairs loop plus a triangle-box prefilter: they must agree
    assert np.array_equal(sc.brute_bits(poses[:60]), sc.brute_bits_fast(poses[:60]))


def _grazing_poses(sc, rng, env, n_seeds, steps=14):
    """Pairs (colliding pose, free pose) with the same rotation, bisected to the contact boundary; returns poses at the
    boundary and a little to either side of it."""
    hitp = scenes.random_poses_near_objects(rng, 3 * n_seeds, env, 1.0, fmt="tq7")
    hit = O.unpack_bits(sc.oracle_bits(hitp, O.POSE_TQ7), len(hitp))
    a = hitp[hit][:n_seeds]
    d = np.stack([rng.uniform(len(a), -1, 1) for _ in range(3)], 1)
    d /= np.linalg.norm(d, axis=1)[:, None]
    b = a.copy()
    b[:, :3] += 14.0 * d
    free = ~O.unpack_bits(sc.oracle_bits(b, O.POSE_TQ7), len(b))
    a, b = a[free], b[free]
    lo_t, hi_t = np.zeros(len(a)), np.ones(len(a))      # collides at lo_t, free at hi_t
    for _ in range(steps):
        mid = 0.5 * (lo_t + hi_t)
        m = a.copy()
        m[:, :3] = a[:, :3] + mid[:, None] * (b[:, :3] - a[:, :3])
        h = O.unpack_bits(sc.oracle_bits(m, O.POSE_TQ7), len(m))
        lo_t = np.where(h, mid, lo_t)
        hi_t = np.where(h, hi_t, mid)
    out = []
    for off in (0.0, 2e-5, -2e-5, 3e-4, -3e-4, 4e-3, -4e-3):
        for t in (lo_t, hi_t):
            m = a.copy()
            m[:, :3] = a[:, :3] + (t + off)[:, None] * (b[:, :3] - a[:, :3])
            out.append(m)
    return np.concatenate(out)


def test_bruteforce_equals_bvh_oracle_on_grazing_c5_poses(built):
    """>= 20 000 poses of the C5 scene concentrated at the collide/free boundary (first/last contact along a ray)."""
    env = scenes.synthetic_environment(2000)
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    sc = Scene(env, [agent])
    rng = scenes.SplitMix64(0xC5C5)
    poses = _grazing_poses(sc, rng, env, 1800)
    assert len(poses) >= 20000
    want = sc.oracle_bits(poses, O.POSE_TQ7)
    frac = O.unpack_bits(want, len(poses)).mean()
    assert 0.25 < frac < 0.75, frac
    got = sc.brute_bits_fast(poses, O.POSE_TQ7)
    bad = np.nonzero(O.unpack_bits(want, len(poses)) != O.unpack_bits(got, len(poses)))[0]
    assert bad.size == 0, "BVH oracle and brute force differ on %d grazing poses, first %s" % (bad.size, bad[:5])


def _orient(a, b, c, d):
    """sign-carrying volume of (a, b, c, d) in extended precision"""
    m = np.stack([b - a, c - a, d - a], axis=-2).astype(np.longdouble)
    return (m[..., 0, 0] * (m[..., 1, 1] * m[..., 2, 2] - m[..., 1, 2] * m[..., 2, 1])
            - m[..., 0, 1] * (m[..., 1, 0] * m[..., 2, 2] - m[..., 1, 2] * m[..., 2, 0])
            + m[..., 0, 2] * (m[..., 1, 0] * m[..., 2, 1] - m[..., 1, 1] * m[..., 2, 0]))


def _segment_hits_triangle(p, q, a, b, c, eps):
    """-> (hit, clear): hit = segment pq crosses triangle abc; clear = no predicate was within eps of zero"""
    s1, s2 = _orient(a, b, c, p), _orient(a, b, c, q)
    t1, t2, t3 = _orient(p, q, a, b), _orient(p, q, b, c), _orient(p, q, c, a)
    clear = (np.abs(s1) > eps) & (np.abs(s2) > eps) & (np.abs(t1) > eps) & (np.abs(t2) > eps) & (np.abs(t3) > eps)
    straddle = (s1 > 0) != (s2 > 0)
    inside = ((t1 > 0) & (t2 > 0) & (t3 > 0)) | ((t1 < 0) & (t2 < 0) & (t3 < 0))
    return straddle & inside, clear


def test_triangle_test_equals_segment_triangle_predicates(built):
    rng = np.random.default_rng(20260)
    n = 30000
    P = rng.normal(size=(n, 3, 3))
    Q = rng.normal(size=(n, 3, 3)) * rng.uniform(0.2, 1.5, size=(n, 1, 1)) + rng.normal(size=(n, 1, 3)) * 0.6
    hit = np.zeros(n, bool)
    clear = np.ones(n, bool)
    for (A, B) in ((P, Q), (Q, P)):
        for i in range(3):
            h, c = _segment_hits_triangle(A[:, i], A[:, (i + 1) % 3], B[:, 0], B[:, 1], B[:, 2], 1e-7)
            hit |= h
            clear &= c
    got = np.array([O.tri_intersect(P[k].ravel(), Q[k].ravel()) for k in np.nonzero(clear)[0]])
    assert clear.sum() > 0.9 * n and 0.15 < hit[clear].mean() < 0.85
    assert np.array_equal(got, hit[clear])
    # the same pairs seen through a rigid motion: Q given in its own frame, (R, T) brings it to P's
    Rm = scenes.quat_to_matrix(scenes.random_quaternions(scenes.SplitMix64(7), n))
    T = rng.normal(size=(n, 3))
    Ql = np.einsum("nji,nkj->nki", Rm, Q - T[:, None, :])           # Ql = R^T (Q - T)
    idx = np.nonzero(clear)[0][:6000]
    got = np.array([O.tri_intersect(P[k].ravel(), Ql[k].ravel(), Rm[k], T[k]) for k in idx])
    assert np.array_equal(got, hit[idx])


def test_obb_test_equals_corner_projection(built):
    """orc_obb_overlap (B-matrix form, 1e-6 pad) vs projecting the 8 + 8 corners on the 15 candidate axes."""
    rng = np.random.default_rng(777)
    sm = scenes.SplitMix64(778)
    n = 20000
    A = scenes.quat_to_matrix(scenes.random_quaternions(sm, n))     # rows = axes of box 1 (its own frame)
    B = scenes.quat_to_matrix(scenes.random_quaternions(sm, n))
    R0 = scenes.quat_to_matrix(scenes.random_quaternions(sm, n))
    ea = np.abs(rng.normal(size=(n, 3))) * np.array([1.0, 0.5, 0.05]) + 1e-3
    eb = np.abs(rng.normal(size=(n, 3))) * np.array([1.0, 0.5, 0.05]) + 1e-3
    ca, cb = 0.3 * rng.normal(size=(n, 3)), 0.3 * rng.normal(size=(n, 3))
    T0 = rng.normal(size=(n, 3)) * 0.45
    sg = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float)
    LD = np.longdouble
    # corners in frame 1: box 1 directly, box 2 through (R0, T0)
    c1 = ca[:, None, :] + np.einsum("cj,nj,nji->nci", sg, ea, A)
    c2l = cb[:, None, :] + np.einsum("cj,nj,nji->nci", sg, eb, B)
    c2 = np.einsum("nij,ncj->nci", R0, c2l) + T0[:, None, :]
    b_axes = np.einsum("nij,nkj->nki", R0, B)                         # rows: R0 b_k
    axes = [A[:, i] for i in range(3)] + [b_axes[:, j] for j in range(3)]
    axes += [np.cross(A[:, i], b_axes[:, j]) for i in range(3) for j in range(3)]
    gap = np.full(n, -np.inf, dtype=LD)
    for L in axes:
        p1 = np.einsum("nci,ni->nc", c1.astype(LD), L.astype(LD))
        p2 = np.einsum("nci,ni->nc", c2.astype(LD), L.astype(LD))
        g = np.maximum(p1.min(1) - p2.max(1), p2.min(1) - p1.max(1))
        gap = np.maximum(gap, g)
    got = np.array([O.obb_overlap(R0[k], T0[k], (A[k], ca[k], ea[k]), (B[k], cb[k], eb[k])) for k in range(n)])
    pad = 1e-5
    sure_overlap, sure_disjoint = gap < -1e-9, gap > pad
    assert sure_overlap.sum() > 1000 and sure_disjoint.sum() > 1000
    assert got[sure_overlap].all(), "FCL's test separated boxes whose projections overlap on all 15 axes"
    assert not got[sure_disjoint].any(), "FCL's test kept boxes that an axis separates by more than its 1e-6 pad"
