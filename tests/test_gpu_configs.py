"""BASELINE.json configs C1-C4 as parity cases: pose batches produced with the reference agents' own
rules (scenes.py restates getPoses / doStep / stateToFCLTransform(s) verbatim), GPU == oracle."""
import numpy as np
import pytest

from helpers import Scene, assert_bits_equal, env_bbox, unit_box
from motionplanningtoolkit_b200 import meshio, scenes
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
