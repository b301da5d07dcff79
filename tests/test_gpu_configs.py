"""BASELINE.json configs C1-C4 as parity cases: pose batches produced with the reference agents' own
rules (scenes.py restates getPoses / doStep / stateToFCLTransform(s) verbatim), GPU == oracle."""
import numpy as np
import pytest

from helpers import Scene, assert_bits_equal, env_bbox, unit_box
from motionplanningtoolkit_b200 import capi, meshio, scenes
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_c1_omnidirectional_rrt_edges(built):
    """omnidirectional.inst: unit_box agent / unit_box environment, unit random steps at dt = 0.1."""
    b = unit_box()
    sc = Scene([b], [b])
    rng = scenes.SplitMix64(scenes.SEED_C1)
    segs, allp = [0], []
    for _ in range(3000):
        s = np.array([rng.uniform(1, -3, 3)[0] for _ in range(3)])
        d = np.array([rng.uniform(1, -1, 1)[0] for _ in range(3)])
        pts = scenes.omnidirectional_edge_poses(s, s + d / np.linalg.norm(d), 0.1)
        allp.append(pts)
        segs.append(segs[-1] + len(pts))
    poses = scenes.translation_poses(np.concatenate(allp))
    chk = sc.checker()
    assert_bits_equal(chk.check_poses(poses), sc.oracle_bits(poses), len(poses), "C1 poses")
    safe = chk.check_segments(segs, poses)
    hit = O.unpack_bits(sc.oracle_bits(poses), len(poses))
    want = np.array([not hit[segs[i]:segs[i + 1]].any() for i in range(len(segs) - 1)])
    assert np.array_equal(safe, want) and 0 < want.sum() < len(want)


def test_c2_prm_roadmap_edges(built):
    """PRM slices: k = 10 nearest earlier milestones, full straight edges at dt = 0.1."""
    b = unit_box()
    sc = Scene([b], [b])
    rng = scenes.SplitMix64(scenes.SEED_C2)
    chk = sc.checker()
    pts = None
    for _ in range(2):
        poses, segs, pts = scenes.prm_slice_edges(rng, 100, (-10,) * 3, (10,) * 3, existing=pts)
        assert_bits_equal(chk.check_poses(poses), sc.oracle_bits(poses), len(poses), "C2 poses")
        safe = chk.check_segments(segs, poses)
        hit = O.unpack_bits(sc.oracle_bits(poses), len(poses))
        want = np.array([not hit[int(segs[i]):int(segs[i + 1])].any() for i in range(len(segs) - 1)])
        assert np.array_equal(safe, want)
    assert (~want).any()


@pytest.mark.parametrize("mode", ["last", "merge"])
def test_c3_blimp_edges(built, mode):
    """blimp.3ds (reference sub-mesh rule and merged) vs divider.dae, doStep-integrated edges, 10 poses each."""
    env = meshio.environment_submeshes(meshio.load_mptm("divider"))
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), mode)
    sc = Scene(env, [agent])
    lo, hi = env_bbox(env, 0.0)
    rng = scenes.SplitMix64(scenes.SEED_C3)
    poses, segs = scenes.blimp_edge_poses(rng, 3000, lo, hi)
    chk = sc.checker()
    want = sc.oracle_bits(poses)
    assert_bits_equal(chk.check_poses(poses), want, len(poses), "C3 " + mode)
    assert O.unpack_bits(want, len(poses)).any()


def test_c4_snake_trailers(built):
    """snake.inst geometry: 11 links of unit_box, stateToFCLTransforms verbatim, obstacle at the origin and
    at (-100, 0, 0) (snake.inst:17)."""
    b = unit_box()
    rng = scenes.SplitMix64(scenes.SEED_C4)
    states = scenes.snake_random_states(rng, 6000, (-4, -4), (4, 4))
    poses = scenes.snake_state_poses(states)
    assert poses.shape == (6000, 11, 12)
    for pose7 in ((0, 0, 0, 1, 0, 0, 0), (-100, 0, 0, 1, 0, 0, 0)):
        sc = Scene([b], [b] * 11, pose7)
        chk = sc.checker()
        want = sc.oracle_bits(poses)
        assert_bits_equal(chk.check_poses(poses), want, len(poses), "C4 %s" % (pose7,))
    # with the obstacle at (-100,0,0) nothing is reachable; at the origin the trailers (all parked at
    # x = -(L+h), Q4) and the tractor collide for a share of the states
    assert not O.unpack_bits(want, len(poses)).any()


def test_non_finite_poses_match_oracle(built):
    """NaN / inf transforms: FCL's comparisons all fail 'open', the oracle defines the verdict."""
    b = unit_box()
    sc = Scene([b], [b])
    rng = scenes.SplitMix64(5)
    poses = scenes.random_poses_rt12(rng, 256, (-1.5,) * 3, (1.5,) * 3)
    poses[3, 9] = np.nan
    poses[40, 0] = np.nan
    poses[77, 10] = np.inf
    poses[130, 4] = -np.inf
    chk = sc.checker()
    assert_bits_equal(chk.check_poses(poses), sc.oracle_bits(poses), len(poses), "non-finite poses")


def test_on_device_edge_interpolation(built):
    """N2: mptcc_check_omni_edges == check_segments over host-interpolated poses == the oracle."""
    b = unit_box()
    sc = Scene([b], [b])
    rng = scenes.SplitMix64(31)
    n = 5000
    start = np.stack([rng.uniform(n, -3, 3) for _ in range(3)], 1)
    length = rng.uniform(n, 0.0, 4.0)
    d = np.stack([rng.uniform(n, -1, 1) for _ in range(3)], 1)
    end = start + d / np.linalg.norm(d, axis=1, keepdims=True) * length[:, None]
    end[:5] = start[:5]                                   # zero-length edges: {start, end}
    end[5:10] = start[5:10] + [[1.0, 0, 0]] * 5           # dist/dt exactly 10: no end pose (reference quirk)
    edges = np.concatenate([start, end], axis=1)
    chk = sc.checker()
    got = chk.check_omni_edges(edges, 0.1)
    segs, allp = [0], []
    for i in range(n):
        pts = scenes.omnidirectional_edge_poses(start[i], end[i], 0.1)
        allp.append(pts)
        segs.append(segs[-1] + len(pts))
    poses = scenes.translation_poses(np.concatenate(allp))
    assert np.array_equal(got, chk.check_segments(segs, poses))
    hit = O.unpack_bits(sc.oracle_bits(poses), len(poses))
    want = np.array([not hit[segs[i]:segs[i + 1]].any() for i in range(n)])
    assert np.array_equal(got, want) and 0 < want.sum() < n


@pytest.mark.parametrize("opts", [dict(split=0), dict(split=1, split_min=32)])
def test_c5_single_object_terrain(built, opts):
    """BASELINE C5, single-object variant: one ~1.0 M-triangle heightfield (depth-23 OBB tree) vs the merged
    blimp; poses hover around the surface so that grazing cases are frequent."""
    env = [scenes.terrain(708, 708)]
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    sc = Scene(env, [agent])
    rng = scenes.SplitMix64(scenes.SEED_C5)
    n = 20000
    poses = scenes.random_poses_rt12(rng, n, (-95, -95, -10), (95, 95, 12))
    chk = sc.checker(**opts)
    want = sc.oracle_bits(poses)
    assert_bits_equal(chk.check_poses(poses), want, n, "terrain %s" % (opts,))
    f = O.unpack_bits(want, n).mean()
    assert 0.1 < f < 0.9, f


def _independent_link_poses(rng, n, n_links, lo, hi):
    """[n, n_links, 12]: every link gets its own random rigid pose (so that some pose-sets self-collide and
    some do not)."""
    return np.stack([scenes.random_poses_rt12(rng, n, lo, hi) for _ in range(n_links)], axis=1)


def test_self_collision_unit_box_links(built):
    """checkSelfCollision = true == oracle (self-collision of the pose-set OR environment)."""
    b = unit_box()
    sc = Scene([b], [b, b, b, b])
    rng = scenes.SplitMix64(0x5E1F)
    poses = _independent_link_poses(rng, 20000, 4, (-2.2,) * 3, (2.2,) * 3)
    # exact contacts: axis-aligned links one unit apart
    T = np.zeros((64, 4, 3))
    T[:, 1, 0] = 1.0
    T[:, 2, 0] = np.linspace(2.0, 2.0 + 2.0 ** -50, 64)
    T[:, 3, 1] = np.linspace(0.9, 1.1, 64)
    T[:, :, 2] += 7.0  # away from the environment box at the origin
    poses = np.concatenate([poses, scenes.translation_poses(T.reshape(-1, 3)).reshape(64, 4, 12)])
    n = len(poses)
    want_self = sc.oracle_bits_self(poses)
    want_env = sc.oracle_bits(poses)
    assert O.unpack_bits(want_self, n).sum() > O.unpack_bits(want_env, n).sum() > 0
    assert not O.unpack_bits(want_self, n).all()
    chk = sc.checker(self_collision=1)
    assert_bits_equal(chk.check_poses(poses), want_self, n, "self-collision on")
    chk.set_option("self_collision", 0)
    assert_bits_equal(chk.check_poses(poses), want_env, n, "self-collision off")
    # quaternion records
    tq = np.stack([scenes.random_poses_tq7(rng, 5000, (-2.0,) * 3, (2.0,) * 3) for _ in range(4)], axis=1)
    chk.set_option("self_collision", 1)
    assert_bits_equal(chk.check_poses(tq, capi.POSE_TQ7), sc.oracle_bits_self(tq, O.POSE_TQ7), len(tq), "self TQ7")


def test_self_collision_mesh_links_and_snake(built):
    """Two blimps as links (deep trees, degenerate triangles -> path replay) near the divider; and the
    reference's snake pose-sets, whose trailers all sit at the same place (always self-colliding)."""
    env = meshio.environment_submeshes(meshio.load_mptm("divider"))
    blimp = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    sc = Scene(env, [blimp, blimp])
    rng = scenes.SplitMix64(0x5E20)
    lo, hi = env_bbox(env, 3.0)
    base = scenes.random_poses_rt12(rng, 3000, lo, hi)
    other = scenes.random_poses_rt12(rng, 3000, (-6.0,) * 3, (6.0,) * 3)  # second blimp: random offset from the first
    second = base.copy()
    second[:, 9:12] += other[:, 9:12]
    second[:, 0:9] = other[:, 0:9]
    poses = np.stack([base, second], axis=1)
    n = len(poses)
    want = sc.oracle_bits_self(poses)
    frac = O.unpack_bits(want, n).mean()
    assert 0.1 < frac < 0.9, frac
    for opts in (dict(self_collision=1), dict(self_collision=1, always_replay=1)):
        assert_bits_equal(sc.checker(**opts).check_poses(poses), want, n, "blimp pair %s" % (opts,))
    b = unit_box()
    snake = Scene([b], [b] * 11, env_pose7=(-100.0, 0, 0, 1, 0, 0, 0))
    sp = scenes.snake_state_poses(scenes.snake_random_states(rng, 2000, (-4, -4), (4, 4)))
    want = snake.oracle_bits_self(sp)
    assert O.unpack_bits(want, 2000).all()
    assert_bits_equal(snake.checker(self_collision=1).check_poses(sp), want, 2000, "snake self")


def test_c5_full_size_properties(built):
    """BASELINE C5 at its full size (2000 objects / 1.14 M triangles, 2^20 pose-sets): size-independent properties
    of the verdict function, plus the oracle on a sample.
      * the two-kernel pipeline, the fused kernel and the forced path replay give the same words;
      * a verdict depends on its own pose-set only: permuting the batch permutes the bits, and checking the
        batch in two halves (different sub-batch cuts, different warps) gives the same bits;
      * the oracle agrees on every 53rd pose-set."""
    env = scenes.synthetic_environment(2000)
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    sc = Scene(env, [agent])
    n = 1 << 20
    rng = scenes.SplitMix64(0x5EED0005)
    lo, hi = env_bbox(env)
    n_near = (n * 2) // 5
    poses = np.concatenate([scenes.random_poses_rt12(rng, n - n_near, lo, hi),
                            scenes.random_poses_near_objects(rng, n_near, env, 3.0)])
    chk = sc.checker()
    split = chk.check_poses(poses)
    bits = O.unpack_bits(split, n)
    assert 0.2 < bits.mean() < 0.5
    chk.set_option("split", 0)
    assert np.array_equal(chk.check_poses(poses), split), "fused kernel != split pipeline"
    chk.set_option("split", 1)
    chk.set_option("always_replay", 1)
    assert np.array_equal(chk.check_poses(poses), split), "path certificate != full replay"
    chk.set_option("always_replay", 0)
    perm = np.argsort(rng.uniform(n, 0.0, 1.0), kind="stable")
    assert np.array_equal(O.unpack_bits(chk.check_poses(poses[perm]), n), bits[perm]), "not permutation-equivariant"
    cut = 419_431  # not a multiple of 32: the halves start on different word / warp boundaries
    two = np.concatenate([O.unpack_bits(chk.check_poses(poses[:cut]), cut),
                          O.unpack_bits(chk.check_poses(poses[cut:]), n - cut)])
    assert np.array_equal(two, bits), "verdicts depend on the batch cut"
    sample = np.arange(0, n, 53)
    want = O.unpack_bits(sc.oracle_bits(poses[sample], n_threads=16), len(sample))
    assert np.array_equal(bits[sample], want), "%d of %d sampled verdicts differ from the oracle" % (
        int((bits[sample] != want).sum()), len(sample))
