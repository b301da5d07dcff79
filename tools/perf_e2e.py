#!/usr/bin/env python3
"""End-to-end rate of blocking mptcc_check_poses calls on the C5 batch (pinned host buffers), for tuning the host
pipeline.  usage: perf_e2e.py [fmt=tq7|rt12] [opt=value ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from motionplanningtoolkit_b200 import capi  # noqa: E402

fmt = "tq7"
opts = []
for a in sys.argv[1:]:
    k, v = a.split("=")
    if k == "fmt":
        fmt = v
    else:
        opts.append((k, int(v)))
n = 1 << 20
env, agent = bench.make_scene()
chk = capi.Checker(1)
chk.set_environment(env)
chk.set_agent([agent])
F = bench.FMT_NAMES[fmt]
hp, hb = [], []
for b in range(4):
    p = bench.make_poses(env, n, 0x5EED0005 + b, fmt)
    a = capi.PinnedArray(p.shape, np.float64)
    a.array[...] = p
    hp.append(a)
    hb.append(capi.PinnedArray(((n + 31) // 32,), np.uint32))
base = None
for label, oo in [("default", [])] + [("%s=%d" % o, [o]) for o in opts]:
    for k, v in oo:
        chk.set_option(k, v)
    for i in range(4):
        chk.check_poses(hp[i].array, F, out=hb[i].array)
    t0 = time.perf_counter()
    reps = 40
    for i in range(reps):
        chk.check_poses(hp[i % 4].array, F, out=hb[i % 4].array)
    dt = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    prev = None
    for i in range(reps):
        t = chk.submit_poses(hp[i % 4].array, hb[i % 4].array, F)
        if prev is not None:
            chk.wait(prev)
        prev = t
    chk.wait(prev)
    dp = (time.perf_counter() - t0) / reps
    h = hash(hb[0].array.tobytes())
    if base is None:
        base = h
    print("%-22s %s  blocking %.3f ms (%.1f M/s)  2-in-flight %.3f ms (%.1f M/s)  device %.3f ms  same_bits %s" % (
        label, fmt, dt * 1e3, n / dt / 1e6, dp * 1e3, n / dp / 1e6, chk.last_kernel_ms(), h == base))
