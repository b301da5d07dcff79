#!/usr/bin/env python3
"""Per-call latency of mptcc_check_poses for planner-sized calls (one edge = ~10 pose-sets), with the CPU path beside
it: the oracle's MeshHandler::isInCollision-equivalent on ONE thread for the same pose-sets (the reference toolkit is
single-threaded).  usage: perf_latency.py [calls] [scene: box|blimp]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from motionplanningtoolkit_b200 import capi, meshio, scenes  # noqa: E402
from oracle import oracle as O  # noqa: E402

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
scene = sys.argv[2] if len(sys.argv) > 2 else "box"
rng = scenes.SplitMix64(7)
if scene == "box":
    box = meshio.load_mptm("unit_box")[0]
    env, agent, lo, hi = [box], box, (-2,) * 3, (2,) * 3
else:
    env = meshio.environment_submeshes(meshio.load_mptm("divider"))
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    lo = np.min([m[0].min(0) for m in env], axis=0)
    hi = np.max([m[0].max(0) for m in env], axis=0)
chk = capi.Checker(1)
chk.set_environment(env)
chk.set_agent([agent])
om = {}
oenv = O.Environment([om.setdefault(id(m), O.Model(*m)) for m in env])
olink = [O.Model(*agent)]
print("scene %s: %d environment objects, agent %d triangles" % (scene, len(env), len(agent[1])))
for n in (1, 10, 100, 500, 1000, 2000, 10000):
    poses = scenes.random_poses_rt12(rng, n, lo, hi)
    out = np.zeros((n + 31) // 32, np.uint32)
    for _ in range(50):
        chk.check_poses(poses, out=out)
    t0 = time.perf_counter()
    for _ in range(calls):
        chk.check_poses(poses, out=out)
    dt = (time.perf_counter() - t0) / calls
    pp, pb = capi.PinnedArray(poses.shape, np.float64), capi.PinnedArray(out.shape, np.uint32)
    pp.array[...] = poses
    for _ in range(50):
        chk.check_poses(pp.array, out=pb.array)
    t0 = time.perf_counter()
    for _ in range(calls):
        chk.check_poses(pp.array, out=pb.array)
    dtp = (time.perf_counter() - t0) / calls
    reps = max(3, min(calls, int(2e5 / n)))
    O.verdicts(oenv, olink, poses, O.POSE_RT12, 1)
    t0 = time.perf_counter()
    for _ in range(reps):
        O.verdicts(oenv, olink, poses, O.POSE_RT12, 1)
    dc = (time.perf_counter() - t0) / reps
    print("n=%6d pose-sets per call: GPU %8.1f us pageable, %8.1f us pinned | CPU oracle 1 thread %9.1f us | GPU/CPU %.2f"
          % (n, dt * 1e6, dtp * 1e6, dc * 1e6, dt / dc))
