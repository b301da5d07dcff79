"""CPU tests of the oracle: analytic known answers (independent of any restatement), BVH invariants, a
structure-free brute-force cross-check, and the committed golden vectors.  The reference has no tests
or golden vectors of its own, so parity with FCL itself is UNPINNED; these tests pin
the oracle's semantics."""
import glob
import os

import numpy as np
import pytest

from helpers import Scene, unit_box
from motionplanningtoolkit_b200 import meshio, scenes
from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def box_scene(built):
    b = unit_box()
    return Scene([b], [b])


def test_unit_cubes_collide_iff_chebyshev_distance_at_most_one(box_scene):
    """Two axis-aligned unit cubes: surfaces meet iff max|d| <= 1; touching counts (FCL's SAT is strict)."""
    e = 2.0 ** -52
    T = np.array([[0, 0, 0], [1, 0, 0], [1 + e, 0, 0], [1 - e, 0, 0], [0, 1, 0], [0, 0, -1], [0, 0, -1 - e],
                  [1, 1, 1], [1, 1, 1 + e], [2, 0, 0], [0.3, 0.99, -1], [-1, -1, -1], [0.5, 0.5, 0.5], [5, 5, 5]], float)
    got = O.unpack_bits(box_scene.oracle_bits(scenes.translation_poses(T), n_threads=1), len(T))
    assert np.array_equal(got, np.max(np.abs(T), axis=1) <= 1.0)
    rng = scenes.SplitMix64(1)
    T = np.stack([rng.uniform(5000, -1.5, 1.5) for _ in range(3)], 1)
    got = O.unpack_bits(box_scene.oracle_bits(scenes.translation_poses(T)), len(T))
    assert np.array_equal(got, np.max(np.abs(T), axis=1) <= 1.0)


def test_containment_is_not_collision(built):
    """Mesh-mesh collision is SURFACE collision: a small box strictly inside a big closed box is free."""
    v, t = unit_box()
    big = (v * 10.0, t)
    sc = Scene([big], [(v, t)])
    T = np.array([[0, 0, 0], [1, 2, -3], [4.4, 0, 0], [4.5, 0, 0], [4.6, 0, 0], [5.5, 0, 0], [5.6, 0, 0]], float)
    got = O.unpack_bits(sc.oracle_bits(scenes.translation_poses(T)), len(T))
    assert list(got) == [False, False, False, True, True, True, False]


def test_bvh_invariants(built):
    for name in ("unit_box", "blimp", "model"):
        for v, t in meshio.environment_submeshes(meshio.load_mptm(name)):
            m = O.Model(v, t)
            nd = m.nodes()
            n = len(t)
            assert m.num_nodes == 2 * n - 1
            leaves = nd["first_child"] < 0
            prims = -(nd["first_child"][leaves] + 1)
            assert sorted(prims.tolist()) == list(range(n))          # every triangle in exactly one leaf
            assert sorted(m.primitive_indices().tolist()) == list(range(n))
            inner = np.nonzero(~leaves)[0]
            fc = nd["first_child"][inner]
            assert np.all(nd["num_primitives"][fc] + nd["num_primitives"][fc + 1] == nd["num_primitives"][inner])
            # every OBB contains the corners of its triangles
            perm = m.primitive_indices()
            for i in range(m.num_nodes):
                idx = perm[nd["first_primitive"][i]:nd["first_primitive"][i] + nd["num_primitives"][i]]
                pts = v[t[idx].reshape(-1)]
                local = (pts - nd["To"][i]) @ nd["axis"][i].T
                assert np.all(np.abs(local) <= nd["extent"][i] * (1 + 1e-12) + 1e-12)


def test_local_aabb_and_object_aabb(built):
    v, t = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    m = O.Model(v, t)
    box, ctr, rad = m.local_aabb()
    assert np.allclose(box[:3], v.min(0)) and np.allclose(box[3:], v.max(0))
    assert np.isclose(rad, np.sqrt(((v - ctr) ** 2).sum(1).max()))
    # identity quaternion: translated tight box; otherwise the sphere box around the moved centre
    rec = np.concatenate([np.eye(3).ravel(), [1.0, 2.0, 3.0]])
    assert np.array_equal(m.object_aabb(rec), np.concatenate([box[:3] + [1, 2, 3], box[3:] + [1, 2, 3]]))
    q = scenes.random_quaternions(scenes.SplitMix64(2), 1)
    R = scenes.quat_to_matrix(q)[0]
    a = m.object_aabb(np.concatenate([R.ravel(), [1.0, 2.0, 3.0]]))
    c = R @ ctr + [1, 2, 3]
    assert np.allclose(a[:3], c - rad) and np.allclose(a[3:], c + rad)


def test_pose_decode_roundtrip(built):
    q = scenes.random_quaternions(scenes.SplitMix64(3), 64)
    for k in range(64):
        R, T, qq = O.pose_decode(np.concatenate([[0.5, -1, 2], q[k]]), O.POSE_TQ7)
        assert np.allclose(R, scenes.quat_to_matrix(q[k:k + 1])[0], atol=1e-15)
        R2, _, q2 = O.pose_decode(np.concatenate([R.ravel(), [0.5, -1, 2]]), O.POSE_RT12)
        assert np.array_equal(R2, R)
        assert np.allclose(np.abs(q2 @ q[k]), 1.0, atol=1e-12)          # same rotation up to sign
    _, _, qi = O.pose_decode(np.concatenate([np.eye(3).ravel(), [0, 0, 0]]), O.POSE_RT12)
    assert list(qi) == [1.0, 0.0, 0.0, 0.0]


def test_bvh_oracle_equals_bruteforce(built):
    """Pruning equivalence: the BVH traversal and the all-pairs triangle test agree."""
    box = unit_box()
    rng = scenes.SplitMix64(9)
    sc = Scene([box], [box])
    poses = scenes.random_poses_rt12(rng, 4000, (-1.8,) * 3, (1.8,) * 3)
    assert np.array_equal(sc.oracle_bits(poses), sc.brute_bits(poses))
    env = meshio.environment_submeshes(meshio.load_mptm("model"))
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "last")
    sc = Scene(env, [agent])
    poses = scenes.random_poses_near_objects(rng, 1500, env, 1.0)
    assert np.array_equal(sc.oracle_bits(poses), sc.brute_bits(poses))


def test_is_in_collision_batch_semantics(box_scene):
    """MeshHandler::isInCollision returns ONE bool for all pose-sets of an edge."""
    far = scenes.translation_poses([[5, 5, 5], [6, 5, 5]])
    hit = scenes.translation_poses([[5, 5, 5], [0.2, 0, 0], [6, 5, 5]])
    assert not O.is_in_collision(box_scene.oenv, box_scene.olinks, far)
    assert O.is_in_collision(box_scene.oenv, box_scene.olinks, hit)
    assert not O.is_in_collision(box_scene.oenv, box_scene.olinks, np.zeros((0, 12)))


def test_threads_do_not_change_verdicts(box_scene):
    rng = scenes.SplitMix64(10)
    poses = scenes.random_poses_rt12(rng, 3001, (-1.5,) * 3, (1.5,) * 3)
    assert np.array_equal(box_scene.oracle_bits(poses, n_threads=1), box_scene.oracle_bits(poses, n_threads=7))


def _golden_scene(name):
    box = unit_box()
    blimp = meshio.load_mptm("blimp")
    if name.startswith("c1_"):
        return [box], [box]
    if name.startswith("c4_"):
        return [box], [box] * 11
    if name.startswith("n4_"):
        return [box], [box] * 4
    agent = meshio.agent_submesh(blimp, "last" if "_last_" in name else "merge")
    env = meshio.environment_submeshes(meshio.load_mptm("divider" if "divider" in name else "model"))
    return env, [agent]


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))), ids=os.path.basename)
def test_golden_vectors(built, path):
    g = np.load(path)
    env, links = _golden_scene(os.path.basename(path))
    sc = Scene(env, links, g["pose7"])
    ctr = O.Counters()
    with_self = "self_collision" in g.files and int(g["self_collision"])
    bits = (sc.oracle_bits_self if with_self else sc.oracle_bits)(g["poses"], int(g["fmt"]), counters=ctr)
    assert np.array_equal(bits, g["bits"])
    assert ctr.n_bv_tests == int(g["n_bv_tests"]) and ctr.n_leaf_tests == int(g["n_leaf_tests"])


def test_golden_pin_status(built, capsys):
    """Every golden file says whether its verdicts were ever compared with REAL FCL (tools/pin_golden_with_fcl.py
    through oracle/_ref/fcl_harness).  In this repository's build container FCL does not exist, so the honest state
    is "unpinned"; the test reports the state and checks that the pinning route is in place."""
    files = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))
    assert files
    states = {}
    for path in files:
        g = np.load(path)
        assert "pinned_by" in g.files, "%s lacks the pinned_by field (regenerate with tools/make_golden.py)" % path
        states[os.path.basename(path)] = str(g["pinned_by"])
    with capsys.disabled():
        for k, v in states.items():
            print("\n  golden %-42s %s" % (k, ("PINNED by " + v) if v else "unpinned (oracle only)"), end="")
        print()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for rel in ("oracle/fcl_pin/fcl_harness.cpp", "oracle/fcl_pin/CMakeLists.txt", "tools/pin_golden_with_fcl.py"):
        assert os.path.exists(os.path.join(root, rel)), rel
    src = open(os.path.join(root, "oracle/fcl_pin/fcl_harness.cpp")).read()
    assert '#include "utilities/meshhandler.hpp"' in src and "MeshHandler::isInCollision(" in src


def test_self_collision_known_answers():
    """checkSelfCollision (meshhandler.hpp:205-219): unit cubes as links; closed-interval contact counts."""
    b = unit_box()
    m = O.Model(*b)
    links = [m, m, m]
    env = O.Environment([m], (100.0, 0, 0, 1, 0, 0, 0))  # far away: only the self test can fire

    def case(T):
        return scenes.translation_poses(np.array(T, float)).reshape(1, 3, 12)

    touching = case([[0, 0, 0], [1.0, 0, 0], [2.5, 0, 0]])      # links 0-1 share a face
    apart = case([[0, 0, 0], [1.5, 0, 0], [3.0, 0, 0]])
    last_two = case([[0, 0, 0], [1.5, 0, 0], [2.4, 0.3, 0]])    # only links 1-2 overlap
    assert O.self_collision(links, touching) and not O.self_collision(links, apart) and O.self_collision(links, last_two)
    batch = np.concatenate([touching, apart, last_two, apart])
    assert list(O.unpack_bits(O.verdicts_self(env, links, batch), 4)) == [True, False, True, False]
    assert not O.unpack_bits(O.verdicts(env, links, batch), 4).any()
    # a single-link agent has nothing to collide with
    assert not O.self_collision([m], scenes.translation_poses(np.zeros((1, 3))).reshape(1, 1, 12))
    # environment collision OR self-collision
    env2 = O.Environment([m], (3.0, 0, 0, 1, 0, 0, 0))           # touches link 2 of `apart` (x = 3.0)
    assert list(O.unpack_bits(O.verdicts_self(env2, links, batch), 4)) == [True, True, True, True]
    assert list(O.unpack_bits(O.verdicts(env2, links, batch), 4)) == [True, True, True, True]
