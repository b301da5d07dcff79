"""GPU parity tests proper: libmptcc (through the C ABI) == CPU oracle, bit for bit, on seeded inputs."""
import numpy as np
import pytest

from helpers import Scene, assert_bits_equal, env_bbox, unit_box
from motionplanningtoolkit_b200 import capi, meshio, scenes
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def box_scene(built):
    b = unit_box()
    return Scene([b], [b])


def test_unit_box_known_answers(box_scene):
    """two axis-aligned unit cubes collide iff max|d| <= 1 (closed)."""
    e = 2.0 ** -52
    T = np.array([[0, 0, 0], [1, 0, 0], [1 + e, 0, 0], [1 - e, 0, 0], [0, 1, 0], [0, 0, -1], [0, 0, -1 - e],
                  [1, 1, 1], [1, 1, 1 + e], [2, 0, 0], [0.3, 0.99, -1], [-1, -1, -1], [0.5, 0.5, 0.5], [5, 5, 5]], float)
    want = np.max(np.abs(T), axis=1) <= 1.0
    chk = box_scene.checker()
    got = O.unpack_bits(chk.check_poses(scenes.translation_poses(T)), len(T))
    assert np.array_equal(got, want)
    assert np.array_equal(O.unpack_bits(box_scene.oracle_bits(scenes.translation_poses(T)), len(T)), want)


def test_unit_box_random(box_scene):
    rng = scenes.SplitMix64(scenes.SEED_C1)
    n = 20000
    poses = np.concatenate([
        scenes.translation_poses(np.stack([rng.uniform(n, -1.6, 1.6) for _ in range(3)], 1)),
        scenes.random_poses_rt12(rng, n, (-1.8, -1.8, -1.8), (1.8, 1.8, 1.8)),
        scenes.random_poses_rt12(rng, 1000, (-10, -10, -10), (10, 10, 10)),
    ])
    chk = box_scene.checker()
    assert_bits_equal(chk.check_poses(poses), box_scene.oracle_bits(poses), len(poses), "unit_box random")


def test_unit_box_near_touching(box_scene):
    """Adversarial: translations within a few ulps of the touching configuration on every axis."""
    rng = scenes.SplitMix64(7)
    n = 4096
    T = np.stack([rng.uniform(n, -1, 1) for _ in range(3)], 1)
    ax = rng.integers(n, 3)
    sign = np.where(rng.integers(n, 2) == 0, -1.0, 1.0)
    ulps = (rng.integers(n, 9) - 4).astype(np.float64)
    T[np.arange(n), ax] = sign * (1.0 + ulps * 2.0 ** -52)
    poses = scenes.translation_poses(T)
    chk = box_scene.checker()
    assert_bits_equal(chk.check_poses(poses), box_scene.oracle_bits(poses), n, "near touching")


@pytest.mark.parametrize("ppw", [1, 4, 32])
@pytest.mark.parametrize("n", [1, 31, 33, 1000])
def test_ragged_sizes_and_poses_per_warp(box_scene, ppw, n):
    rng = scenes.SplitMix64(100 + n)
    poses = scenes.random_poses_rt12(rng, n, (-1.5, -1.5, -1.5), (1.5, 1.5, 1.5))
    chk = box_scene.checker(poses_per_warp=ppw)
    assert_bits_equal(chk.check_poses(poses), box_scene.oracle_bits(poses), n, "ppw=%d n=%d" % (ppw, n))


def test_empty_batch(box_scene):
    chk = box_scene.checker()
    assert chk.check_poses(np.zeros((0, 12))).size == 0


def test_quaternion_pose_format(box_scene):
    rng = scenes.SplitMix64(11)
    poses = scenes.random_poses_tq7(rng, 5000, (-1.6, -1.6, -1.6), (1.6, 1.6, 1.6))
    poses[:10, 3:] = (1, 0, 0, 0)      # exact identity quaternion -> tight AABB branch
    poses[10:20, 3:] = (-1, 0, 0, 0)   # same rotation, not "identity" for FCL -> sphere AABB branch
    chk = box_scene.checker()
    assert_bits_equal(chk.check_poses(poses, capi.POSE_TQ7), box_scene.oracle_bits(poses, O.POSE_TQ7), len(poses), "TQ7")


def test_blimp_vs_divider(built):
    """C3 geometry: blimp (last sub-mesh, as the reference loads it) vs divider.dae (110 objects)."""
    env = meshio.environment_submeshes(meshio.load_mptm("divider"))
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "last")
    sc = Scene(env, [agent])
    lo, hi = env_bbox(env, 2.0)
    rng = scenes.SplitMix64(scenes.SEED_C3)
    poses = np.concatenate([scenes.random_poses_rt12(rng, 15000, lo, hi),
                            scenes.random_poses_near_objects(rng, 15000, env, 1.0)])
    chk = sc.checker()
    got = chk.check_poses(poses)
    want = sc.oracle_bits(poses)
    assert_bits_equal(got, want, len(poses), "blimp(last) vs divider")
    frac = O.unpack_bits(want, len(poses)).mean()
    assert 0.01 < frac < 0.99, frac


def test_blimp_merged_vs_model_and_divider(built):
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    for name, n in (("model", 6000), ("divider", 6000)):
        env = meshio.environment_submeshes(meshio.load_mptm(name))
        sc = Scene(env, [agent])
        lo, hi = env_bbox(env, 3.0)
        rng = scenes.SplitMix64(scenes.SEED_C3 + 1)
        poses = np.concatenate([scenes.random_poses_rt12(rng, n // 2, lo, hi),
                                scenes.random_poses_near_objects(rng, n - n // 2, env, 2.0)])
        chk = sc.checker()
        assert_bits_equal(chk.check_poses(poses), sc.oracle_bits(poses), n, "blimp(merge) vs " + name)


def test_environment_pose_with_rotation(built):
    """Environment Location with a non-identity quaternion: tf1 enters relativeTransform and the
    sphere-AABB branch of CollisionObject::computeAABB."""
    env = meshio.environment_submeshes(meshio.load_mptm("model"))
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "last")
    q = np.array([0.8, 0.2, -0.4, 0.4])
    q /= np.linalg.norm(q)
    pose7 = np.concatenate([[10.0, -20.0, 5.0], q])
    sc = Scene(env, [agent], pose7)
    rng = scenes.SplitMix64(21)
    poses = scenes.random_poses_rt12(rng, 8000, (-120, -200, -120), (200, 120, 200))
    chk = sc.checker()
    want = sc.oracle_bits(poses)
    assert_bits_equal(chk.check_poses(poses), want, len(poses), "rotated environment")
    assert O.unpack_bits(want, len(poses)).any()


def test_snake_multi_link(built):
    """C4 geometry: 11 links (the same unit_box model) per pose-set."""
    b = unit_box()
    sc = Scene([b], [b] * 11)
    rng = scenes.SplitMix64(scenes.SEED_C4)
    n = 4000
    poses = np.stack([scenes.random_poses_rt12(rng, n, (-4, -4, -1), (4, 4, 1)) for _ in range(11)], axis=1)
    chk = sc.checker()
    want = sc.oracle_bits(poses)
    assert_bits_equal(chk.check_poses(poses), want, n, "snake 11 links")
    f = O.unpack_bits(want, n).mean()
    assert 0.05 < f < 0.99


def test_segments_match_safe_edge_semantics(box_scene):
    rng = scenes.SplitMix64(5)
    segs, allp = [0], []
    for _ in range(300):
        s = np.array([rng.uniform(1, -3, 3)[0] for _ in range(3)])
        e = s + np.array([rng.uniform(1, -1, 1)[0] for _ in range(3)])
        pts = scenes.omnidirectional_edge_poses(s, e, 0.1)
        allp.append(pts)
        segs.append(segs[-1] + len(pts))
    segs.append(segs[-1])  # one empty segment: safe
    poses = scenes.translation_poses(np.concatenate(allp))
    chk = box_scene.checker()
    safe = chk.check_segments(segs, poses)
    for i in range(len(segs) - 1):
        want = not O.is_in_collision(box_scene.oenv, box_scene.olinks, poses[segs[i]:segs[i + 1]]) \
            if segs[i + 1] > segs[i] else True
        assert bool(safe[i]) == want, i


def test_counters_and_exact_path(box_scene):
    rng = scenes.SplitMix64(3)
    poses = scenes.random_poses_rt12(rng, 4096, (-1.2, -1.2, -1.2), (1.2, 1.2, 1.2))
    chk = box_scene.checker(always_replay=1)
    chk.enable_counters(True)
    bits = chk.check_poses(poses)
    c = chk.counters()
    assert_bits_equal(bits, box_scene.oracle_bits(poses), len(poses), "counting kernel")
    assert c["n_pose_sets"] == 4096 and c["n_bv_tests"] > 0 and c["n_leaf_tests"] > 0
    assert c["n_exact_tri"] >= int(O.unpack_bits(bits, len(poses)).sum())
    # with always_replay every positive triangle pair replays FCL's root-to-leaf OBB tests in fp64
    assert c["n_path_checks"] >= int(O.unpack_bits(bits, len(poses)).sum())


def test_path_certificate_equals_full_replay(built):
    """The certificate that skips the fp64 path replay must never change a verdict."""
    env = meshio.environment_submeshes(meshio.load_mptm("divider"))
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    sc = Scene(env, [agent])
    rng = scenes.SplitMix64(77)
    poses = scenes.random_poses_near_objects(rng, 20000, env, 2.0)
    a = sc.checker(always_replay=1)
    b = sc.checker(always_replay=0)
    a.enable_counters(True)
    b.enable_counters(True)
    ba, bb = a.check_poses(poses), b.check_poses(poses)
    assert np.array_equal(ba, bb)
    assert_bits_equal(bb, sc.oracle_bits(poses), len(poses), "certificate")
    assert a.counters()["n_path_checks"] > 0
    assert b.counters()["n_path_checks"] <= a.counters()["n_path_checks"]


import glob  # noqa: E402
import os  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))), ids=os.path.basename)
def test_gpu_matches_golden_vectors(built, path):
    """Committed fixtures (tools/make_golden.py): seeded poses + the oracle's verdict words."""
    from test_oracle_kat import _golden_scene
    g = np.load(path)
    env, links = _golden_scene(os.path.basename(path))
    chk = capi.Checker(1)
    chk.set_environment(env, g["pose7"])
    chk.set_agent(links)
    if "self_collision" in g.files and int(g["self_collision"]):
        chk.set_option("self_collision", 1)
    got = chk.check_poses(g["poses"], int(g["fmt"]))
    assert_bits_equal(got, g["bits"], g["poses"].shape[0], os.path.basename(path))


def test_multi_device_context_matches_single(built):
    """mptcc_create(n_devices=2): contiguous word-aligned slices per device, same verdict words."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = meshio.environment_submeshes(meshio.load_mptm("divider"))
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    sc = Scene(env, [agent])
    rng = scenes.SplitMix64(91)
    poses = scenes.random_poses_near_objects(rng, 50001, env, 2.0)
    one = sc.checker(1).check_poses(poses)
    two = sc.checker(2, sub_batch=4096).check_poses(poses)
    assert np.array_equal(one, two)
    assert_bits_equal(two, sc.oracle_bits(poses), len(poses), "2 devices")


def test_sub_batch_pipeline(box_scene):
    """Host-buffer path: many ping-pong sub-batches give the same words as one launch."""
    rng = scenes.SplitMix64(92)
    poses = scenes.random_poses_rt12(rng, 20011, (-1.6,) * 3, (1.6,) * 3)
    a = box_scene.checker().check_poses(poses)
    b = box_scene.checker(sub_batch=1024).check_poses(poses)
    assert np.array_equal(a, b)
    assert_bits_equal(b, box_scene.oracle_bits(poses), len(poses), "sub-batches")


@pytest.mark.parametrize("opts", [dict(split=0), dict(split=1, split_min=32), dict(split=1, split_min=32, split_cap=1),
                                  dict(split=1, split_min=32, always_replay=1),
                                  dict(split=1, split_min=32, long_batches=1), dict(split=1, split_min=32, long_batches=0)])
def test_split_pipeline_matches_fused_and_oracle(built, opts):
    """Two-kernel pipeline (broadphase -> candidate list -> narrowphase) incl. the overflow fall-back
    (split_cap=1 makes the candidate list too small for this scene)."""
    env = meshio.environment_submeshes(meshio.load_mptm("divider"))
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    sc = Scene(env, [agent])
    rng = scenes.SplitMix64(123)
    poses = scenes.random_poses_near_objects(rng, 30011, env, 2.0)
    chk = sc.checker(**opts)
    chk.enable_counters(True)
    got = chk.check_poses(poses)
    assert_bits_equal(got, sc.oracle_bits(poses), len(poses), "split %s" % (opts,))
    c = chk.counters()
    assert c["n_candidates"] > len(poses) * (1 if opts.get("split_cap") == 1 else 0)


def test_split_pipeline_unit_box_and_quaternions(box_scene):
    rng = scenes.SplitMix64(124)
    poses = scenes.random_poses_tq7(rng, 50000, (-1.6,) * 3, (1.6,) * 3)
    chk = box_scene.checker(split=1, split_min=32)
    assert_bits_equal(chk.check_poses(poses, capi.POSE_TQ7), box_scene.oracle_bits(poses, O.POSE_TQ7), len(poses), "split TQ7")
    poses = scenes.translation_poses(np.stack([rng.uniform(40000, -1.6, 1.6) for _ in range(3)], 1))
    assert_bits_equal(chk.check_poses(poses), box_scene.oracle_bits(poses), len(poses), "split translations")


def test_submit_wait_matches_synchronous_calls(box_scene):
    """mptcc_submit_poses / mptcc_wait: several batches in flight (more than the library keeps outstanding, so a
    submit has to retire the oldest itself), pinned and pageable buffers, tickets waited out of order."""
    rng = scenes.SplitMix64(93)
    chk = box_scene.checker(sub_batch=4096, split=1, split_min=32)
    batches, outs, pinned = [], [], []
    for i in range(7):
        n = 9000 + 517 * i
        poses = scenes.random_poses_rt12(rng, n, (-1.6,) * 3, (1.6,) * 3)
        if i % 2 == 0:  # pinned buffers: the copies really are asynchronous
            pp, pb = capi.PinnedArray(poses.size, np.float64), capi.PinnedArray((n + 31) // 32, np.uint32)
            pp.array[:] = poses.ravel()
            pinned += [pp, pb]
            batches.append(pp.array)
            outs.append(pb.array)
        else:
            batches.append(np.ascontiguousarray(poses).ravel())
            outs.append(np.zeros((n + 31) // 32, np.uint32))
    tickets = [chk.submit_poses(b, o) for b, o in zip(batches, outs)]
    assert tickets == sorted(tickets) and len(set(tickets)) == len(tickets)
    chk.wait(tickets[3])
    chk.wait(tickets[1])  # already complete: returns at once
    chk.wait(0)
    for i, (b, o) in enumerate(zip(batches, outs)):
        n = 9000 + 517 * i
        want = box_scene.oracle_bits(b.reshape(n, 12))
        assert_bits_equal(o, want, n, "submit/wait batch %d" % i)
        assert np.array_equal(o, chk.check_poses(b.reshape(n, 12)))
    with pytest.raises(capi.MptccError):
        chk.wait(tickets[-1] + 100)  # never issued
    assert chk.submit_poses(np.zeros(0), np.zeros(0, np.uint32)) == tickets[-1] + 7  # empty: nothing new to wait for


@pytest.mark.parametrize("n", [1000, 40000])
@pytest.mark.parametrize("offset_doubles", [0, 1, 3])
def test_device_resident_buffers_any_alignment(box_scene, n, offset_doubles):
    """mptcc_check_poses_device: poses already in HBM, on the caller's stream; the 256-bit record loads are only
    taken for 32-byte aligned buffers, every other alignment goes through the scalar loads."""
    import torch
    rng = scenes.SplitMix64(94 + offset_doubles)
    poses = scenes.random_poses_rt12(rng, n, (-1.6,) * 3, (1.6,) * 3)
    want = box_scene.oracle_bits(poses)
    chk = box_scene.checker(split_min=32768)
    buf = torch.zeros(poses.size + 8, dtype=torch.float64, device="cuda")
    buf[offset_doubles:offset_doubles + poses.size] = torch.from_numpy(poses.ravel()).cuda()
    bits = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        chk.check_poses_device(n, buf.data_ptr() + 8 * offset_doubles, bits.data_ptr(), capi.POSE_RT12, 0, stream.cuda_stream)
    stream.synchronize()
    got = bits.cpu().numpy().view(np.uint32)
    assert_bits_equal(got, want, n, "device buffers, offset %d doubles" % offset_doubles)


@pytest.mark.parametrize("opts", [dict(np2=1), dict(np2=1, qnode=1), dict(np2=1, always_replay=1), dict(np2=1, qnode=1, agent_top=0),
                                  dict(np2=1, split_cap=1), dict(np2=1, long_batches=0), dict(np2=1, l2_persist=0)])
def test_second_generation_narrowphase_matches_oracle(built, opts):
    """np2_kernels.cu (TMA-staged agent tree top, child-frame SAT, exact-pass kernel; optionally the 32-byte
    quantised nodes) gives the oracle's verdicts on deep trees, incl. forced path replay and the overflow fall-back."""
    env = meshio.environment_submeshes(meshio.load_mptm("divider"))
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    sc = Scene(env, [agent])
    rng = scenes.SplitMix64(321)
    poses = scenes.random_poses_near_objects(rng, 40007, env, 2.0)
    chk = sc.checker(split=1, split_min=32, **opts)
    assert_bits_equal(chk.check_poses(poses), sc.oracle_bits(poses), len(poses), "np2 %s" % (opts,))
    tq = scenes.random_poses_near_objects(rng, 20000, env, 2.0, fmt="tq7")
    assert_bits_equal(chk.check_poses(tq, capi.POSE_TQ7), sc.oracle_bits(tq, O.POSE_TQ7), len(tq), "np2 TQ7 %s" % (opts,))


@pytest.mark.parametrize("qnode", [0, 1])
def test_second_generation_narrowphase_unit_boxes_and_non_rigid_poses(box_scene, qnode):
    """Exact-contact cases (touching cubes) and poses whose 3x3 block is not a rotation: the np2 kernel evaluates
    agent-side tests with the inverse transform, which only exists for rigid poses, so such a launch must fall
    back to the fused kernel and still give FCL's (formula-defined) verdicts."""
    rng = scenes.SplitMix64(322)
    chk = box_scene.checker(split=1, split_min=32, np2=1, qnode=qnode)
    e = 2.0 ** -52
    T = np.concatenate([np.array([[1, 0, 0], [1 + e, 0, 0], [1 - e, 0, 0], [0, -1, 0], [1, 1, 1], [1, 1, 1 + e]], float),
                        np.stack([rng.uniform(40000, -1.6, 1.6) for _ in range(3)], 1)])
    poses = scenes.translation_poses(T)
    assert_bits_equal(chk.check_poses(poses), box_scene.oracle_bits(poses), len(poses), "np2 translations")
    poses = scenes.random_poses_rt12(rng, 40000, (-1.6,) * 3, (1.6,) * 3)
    poses[::7, :9] *= 1.25                      # scaled "rotations"
    poses[5::11, 1] += 0.3                      # sheared
    assert_bits_equal(chk.check_poses(poses), box_scene.oracle_bits(poses), len(poses), "np2 non-rigid poses")


@pytest.mark.parametrize("opts", [dict(split=0), dict(split=1, split_min=32), dict(split=1, split_min=32, np2=1)])
def test_compact_pose_formats_equal_rt12(built, opts):
    """MPTCC_POSE_T3 (x y z = Transform3f(Vec3f)) and MPTCC_POSE_XYZCS (x y z cos sin = the planar Transform3f(R,T) of
    the blimp and the snake) are the RT12 records with the same entries: identical verdicts, 24 / 40 bytes per pose."""
    env = meshio.environment_submeshes(meshio.load_mptm("divider"))
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    sc = Scene(env, [agent])
    rng = scenes.SplitMix64(555)
    lo, hi = env_bbox(env, 2.0)
    chk = sc.checker(**opts)
    T = np.stack([rng.uniform(36001, lo[k], hi[k]) for k in range(3)], 1)
    want = sc.oracle_bits(scenes.translation_poses(T))
    assert_bits_equal(chk.check_poses(T, capi.POSE_T3), want, len(T), "T3 %s" % (opts,))
    th = rng.uniform(len(T), -np.pi, np.pi)
    xyzcs = np.concatenate([T, np.cos(th)[:, None], np.sin(th)[:, None]], axis=1)
    want = sc.oracle_bits(scenes._rz_convention(th, T))
    assert_bits_equal(chk.check_poses(xyzcs, capi.POSE_XYZCS), want, len(T), "XYZCS %s" % (opts,))
    # planner-sized call (zero-copy path) and multi-link pose-sets
    assert_bits_equal(chk.check_poses(xyzcs[:10], capi.POSE_XYZCS), sc.oracle_bits(scenes._rz_convention(th[:10], T[:10])), 10, "XYZCS small")
    box = meshio.load_mptm("unit_box")[0]
    sc2 = Scene([box], [box] * 3)
    chk2 = sc2.checker(**{k: v for k, v in opts.items() if k != "np2"})
    T3 = np.stack([rng.uniform(3 * 9000, -1.7, 1.7) for _ in range(3)], 1)
    want = sc2.oracle_bits(scenes.translation_poses(T3).reshape(-1, 3, 12))
    assert_bits_equal(chk2.check_poses(T3.reshape(-1, 3, 3), capi.POSE_T3), want, 9000, "T3 3 links")


def test_multi_device_edges_and_device_buffer_overflow(box_scene):
    """mptcc_check_omni_edges on a 2-device context (every device gets its sub-batch before any is waited for) equals
    the single-device result; the device-buffer entry point with a candidate list that is too small falls back and
    still gives the oracle's words, and mptcc_synchronize notes the overflow so the next launch has a longer list."""
    import torch
    rng = scenes.SplitMix64(777)
    n = 300000                                            # > 2^21 poses in total: several sub-batches per device
    start = np.stack([rng.uniform(n, -3, 3) for _ in range(3)], 1)
    d = np.stack([rng.uniform(n, -1, 1) for _ in range(3)], 1)
    end = start + d / np.linalg.norm(d, axis=1, keepdims=True) * rng.uniform(n, 0.0, 2.5)[:, None]
    edges = np.concatenate([start, end], axis=1)
    one = box_scene.checker().check_omni_edges(edges, 0.1)
    assert 0 < one.sum() < n
    if torch.cuda.device_count() >= 2:
        assert np.array_equal(box_scene.checker(2).check_omni_edges(edges, 0.1), one)
    env = meshio.environment_submeshes(meshio.load_mptm("divider"))
    sc = Scene(env, [meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")])
    poses = scenes.random_poses_near_objects(rng, 40000, env, 2.0)
    want = sc.oracle_bits(poses)
    chk = sc.checker(split=1, split_min=32, split_cap=1)
    buf = torch.from_numpy(poses).cuda()
    bits = torch.zeros((len(poses) + 31) // 32, dtype=torch.int32, device="cuda")
    for _ in range(3):                                    # 1st: overflow + fall-back; later: longer list, no overflow
        bits.zero_()
        chk.check_poses_device(len(poses), buf.data_ptr(), bits.data_ptr(), capi.POSE_RT12, 0, None)
        chk.synchronize()
        assert_bits_equal(bits.cpu().numpy().view(np.uint32), want, len(poses), "device buffers, small candidate list")


@pytest.mark.parametrize("opts", [dict(), dict(qnode=0), dict(qnode=0, agent_top=511), dict(always_replay=1), dict(split_cap=1),
                                  dict(grid=0), dict(grid=0, qnode=0), dict(near_first=0), dict(no_steal=1)])
def test_third_generation_narrowphase_matches_oracle(built, opts):
    """np3_kernels.cu, the default narrowphase (streaming lanes, cp.async-prefetched candidate ring, two-wide SAT,
    32-byte quaternion nodes) behind the grid broadphase, on deep trees: every node format (quantised / fp32 / fp32
    with the TMA-staged agent tree top), forced path replay, the overflow fall-back, the cluster-tree broadphase."""
    env = meshio.environment_submeshes(meshio.load_mptm("divider"))
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    sc = Scene(env, [agent])
    rng = scenes.SplitMix64(521)
    poses = scenes.random_poses_near_objects(rng, 40007, env, 2.0)
    chk = sc.checker(split=1, split_min=32, **opts)
    chk.enable_counters(True)
    assert_bits_equal(chk.check_poses(poses), sc.oracle_bits(poses), len(poses), "np3 %s" % (opts,))
    assert chk.counters()["n_candidates"] > 0      # the split pipeline ran (the fused kernel does not count candidates per list)
    tq = scenes.random_poses_near_objects(rng, 20000, env, 2.0, fmt="tq7")
    assert_bits_equal(chk.check_poses(tq, capi.POSE_TQ7), sc.oracle_bits(tq, O.POSE_TQ7), len(tq), "np3 TQ7 %s" % (opts,))


@pytest.mark.parametrize("opts", [dict(), dict(qnode=0), dict(grid=0)])
def test_third_generation_narrowphase_unit_boxes_and_odd_poses(box_scene, opts):
    """Exact-contact cases (touching cubes, +-1 ulp), poses whose 3x3 block is not a rotation (the launch must go to the
    fused kernel) and poses with NaN / inf / far-away translations (the grid broadphase hands such launches to the fused
    kernel as well, which keeps FCL's comparison semantics for them)."""
    rng = scenes.SplitMix64(522)
    chk = box_scene.checker(split=1, split_min=32, **opts)
    e = 2.0 ** -52
    T = np.concatenate([np.array([[1, 0, 0], [1 + e, 0, 0], [1 - e, 0, 0], [0, -1, 0], [1, 1, 1], [1, 1, 1 + e]], float),
                        np.stack([rng.uniform(40000, -1.6, 1.6) for _ in range(3)], 1)])
    poses = scenes.translation_poses(T)
    assert_bits_equal(chk.check_poses(poses), box_scene.oracle_bits(poses), len(poses), "np3 translations %s" % (opts,))
    poses = scenes.random_poses_rt12(rng, 40000, (-1.6,) * 3, (1.6,) * 3)
    poses[::7, :9] *= 1.25                      # scaled "rotations"
    poses[5::11, 1] += 0.3                      # sheared
    assert_bits_equal(chk.check_poses(poses), box_scene.oracle_bits(poses), len(poses), "np3 non-rigid poses %s" % (opts,))
    poses = scenes.random_poses_rt12(rng, 40000, (-1.6,) * 3, (1.6,) * 3)
    poses[3, 9] = np.nan
    poses[1001, 10] = np.inf
    poses[2002, 11] = -np.inf
    poses[3003, 9:12] = (1e30, -1e30, 1e30)
    poses[4004, 9:12] = (1e300, 0.0, 0.0)
    assert_bits_equal(chk.check_poses(poses), box_scene.oracle_bits(poses), len(poses), "np3 non-finite poses %s" % (opts,))


def test_grid_broadphase_candidates_equal_tree_broadphase(built):
    """The grid broadphase and the cluster-tree broadphase both reduce to FCL's exact fp64 AABB::overlap: the same
    number of candidates on a scene of many objects (C5-like, scaled down), incl. poses outside the scene box."""
    env = scenes.synthetic_environment(300, extent=60.0)
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    sc = Scene(env, [agent])
    rng = scenes.SplitMix64(523)
    poses = np.concatenate([scenes.random_poses_tq7(rng, 30000, (-45.0,) * 3, (45.0,) * 3),
                            scenes.random_poses_near_objects(rng, 30000, env, 3.0, fmt="tq7")])
    counts, bits = [], []
    for grid in (1, 0):
        chk = sc.checker(split=1, split_min=32, grid=grid)
        chk.enable_counters(True)
        bits.append(chk.check_poses(poses, capi.POSE_TQ7))
        counts.append(chk.counters()["n_candidates"])
    assert counts[0] == counts[1] and counts[0] > 0
    assert np.array_equal(bits[0], bits[1])
    assert_bits_equal(bits[0], sc.oracle_bits(poses, O.POSE_TQ7), len(poses), "grid broadphase")


def test_quantised_nodes_and_grid_are_in_use_and_fall_back_when_they_do_not_apply(built):
    """The fast structures are really what the default pipeline runs on the fixtures (quantised nodes, grid), and scenes
    they cannot represent -- extents beyond binary16, coordinates beyond what the fp32 cell function resolves -- take the
    fp32 nodes / the cluster tree and still give the oracle's verdicts."""
    env = meshio.environment_submeshes(meshio.load_mptm("divider"))
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    sc = Scene(env, [agent])
    chk = sc.checker(split=1, split_min=32)
    assert chk.info("qnode_active") == 1 and chk.info("grid_cells") > 0 and chk.info("grid_entries") > 0
    rng = scenes.SplitMix64(524)
    # the same scene a hundred thousand times larger: extents up to ~1e6 (no binary16), coordinates ~1e6
    k = 1.0e5
    env_big = [(v * k, t) for v, t in env]
    agent_big = (agent[0] * k, agent[1])
    sb = Scene(env_big, [agent_big])
    chkb = sb.checker(split=1, split_min=32)
    assert chkb.info("qnode_active") == 0
    poses = scenes.random_poses_near_objects(rng, 20011, env_big, 2.0 * k, fmt="tq7")
    assert_bits_equal(chkb.check_poses(poses, capi.POSE_TQ7), sb.oracle_bits(poses, O.POSE_TQ7), len(poses), "scaled scene")
    # a small agent far from the origin: a pose box is ~1 wide at coordinates ~1e7 (fp32 spacing 1): the grid must cope
    far = np.array([1.0e7, -1.0e7, 1.0e7])
    box = unit_box()
    env_far = [(box[0] * 3.0 + far, box[1]), (box[0] * 3.0 + far + np.array([9.0, 0.0, 0.0]), box[1])]
    sf = Scene(env_far, [box])
    chkf = sf.checker(split=1, split_min=32)
    T = far + np.stack([rng.uniform(30000, -8.0, 16.0), rng.uniform(30000, -8.0, 8.0), rng.uniform(30000, -8.0, 8.0)], 1)
    poses = np.concatenate([T, scenes.random_quaternions(rng, len(T))], axis=1)
    assert_bits_equal(chkf.check_poses(poses, capi.POSE_TQ7), sf.oracle_bits(poses, O.POSE_TQ7), len(poses), "far scene")
