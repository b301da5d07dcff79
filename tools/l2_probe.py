#!/usr/bin/env python3
"""Narrowphase time per fp32 OBB test as the environment grows past L2 (same object density, n_objects varies).
usage: l2_probe.py [n_poses]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from motionplanningtoolkit_b200 import capi, meshio, scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
for n_obj in (250, 500, 1000, 2000, 4000):
    env = scenes.synthetic_environment(n_obj, extent=200.0 * (n_obj / 2000.0) ** (1.0 / 3.0))
    poses = bench.make_poses(env, n, 0x5EED0005)
    chk = capi.Checker(1)
    chk.set_environment(env)
    chk.set_agent([agent])
    for kv in sys.argv[2:]:
        k, v = kv.split("=")
        chk.set_option(k, int(v))
    pp = capi.PinnedArray(poses.size, np.float64)
    pp.array[:] = poses.ravel()
    import torch
    d_poses = torch.from_numpy(poses).cuda()
    d_bits = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()
    chk.set_option("time_phases", 1)
    acc = np.zeros(3)
    for i in range(6):
        chk.check_poses_device(n, d_poses.data_ptr(), d_bits.data_ptr(), capi.POSE_RT12, 0, st.cuda_stream)
        torch.cuda.synchronize()
        if i >= 2:
            acc += np.array(chk.last_phase_ms(0))
    acc /= 4
    chk.set_option("time_phases", 0)
    chk.enable_counters(True)
    chk.reset_counters()
    chk.check_poses_device(n, d_poses.data_ptr(), d_bits.data_ptr(), capi.POSE_RT12, 0, st.cuda_stream)
    torch.cuda.synchronize()
    c = chk.counters()
    tris = sum(len(m[1]) for m in env)
    print("objects %5d  node32 %6.1f MB  bp %.3f ms  np %.3f ms  cand/pose %.2f  bv/pose %5.1f  ns per 1000 tests %.1f"
          % (n_obj, 2 * tris * 64 / 1e6, acc[0], acc[1], c["n_candidates"] / n, c["n_bv_tests"] / n,
             1e6 * acc[1] / max(1, c["n_bv_tests"]) * 1000))
    del chk
