#!/usr/bin/env python3
"""Per-call latency of mptcc_check_poses for planner-sized calls (one edge = ~10 pose-sets).
usage: perf_latency.py [calls]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from motionplanningtoolkit_b200 import capi, meshio, scenes  # noqa: E402

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
box = meshio.load_mptm("unit_box")[0]
rng = scenes.SplitMix64(7)
chk = capi.Checker(1)
chk.set_environment([box])
chk.set_agent([box])
for n in (1, 10, 100, 500, 1000, 2000, 10000):
    poses = scenes.random_poses_rt12(rng, n, (-2,) * 3, (2,) * 3)
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
    print("n=%6d pose-sets per call: %8.1f us per call pageable, %8.1f us pinned  (%.2f us per pose-set)"
          % (n, dt * 1e6, dtp * 1e6, dt * 1e6 / n))
