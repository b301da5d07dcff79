#!/usr/bin/env python3
"""Time mptcc_check_omni_edges (poses interpolated on the device) against the same edges expanded on the host
and sent through mptcc_check_poses.  usage: perf_edges.py [n_edges] [edge_length] [dt]
Scene: the C5 synthetic environment, unit-box agent (the Omnidirectional agent of omnidirectional.inst)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from motionplanningtoolkit_b200 import capi, meshio, scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
length = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
dt = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
rng = scenes.SplitMix64(0xED6E)
env = scenes.synthetic_environment(2000)
box = meshio.load_mptm("unit_box")[0]
start = np.stack([rng.uniform(n, -100, 100) for _ in range(3)], 1)
d = np.stack([rng.uniform(n, -1, 1) for _ in range(3)], 1)
d /= np.linalg.norm(d, axis=1, keepdims=True)
edges = np.ascontiguousarray(np.concatenate([start, start + length * d], 1))
chk = capi.Checker(1)
chk.set_environment(env)
chk.set_agent([box])
for _ in range(2):
    safe = chk.check_omni_edges(edges, dt)
t0 = time.perf_counter()
reps = 3
for _ in range(reps):
    safe = chk.check_omni_edges(edges, dt)
t_dev = (time.perf_counter() - t0) / reps
# host expansion of a slice (the reference's getPoses rule) for comparison and a parity spot check
m = min(n, 1 << 16)
t0 = time.perf_counter()
pts = [scenes.omnidirectional_edge_poses(edges[i, :3], edges[i, 3:], dt) for i in range(m)]
t_expand = time.perf_counter() - t0
seg = np.concatenate([[0], np.cumsum([len(p) for p in pts])]).astype(np.uint64)
poses = scenes.translation_poses(np.concatenate(pts))
safe_host = chk.check_segments(seg, poses)
n_poses = int(seg[-1]) * n // m
print("edges %d (%.1f poses/edge)  on-device: %.3f ms = %.1f M edges/s = %.1f M poses/s, %d B/edge over PCIe; unsafe %.3f; "
      "host-expanded slice equal: %s (python expansion of %d edges took %.2f s)"
      % (n, int(seg[-1]) / m, 1e3 * t_dev, n / t_dev / 1e6, n_poses / t_dev / 1e6, 49, 1.0 - safe.mean(),
         bool(np.array_equal(safe[:m], safe_host)), m, t_expand))
