#!/usr/bin/env python3
"""Print libmptcc's traversal counters next to the oracle's FCL-order counts for the C5 workload sample."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from motionplanningtoolkit_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
env, agent = bench.make_scene()
poses = bench.make_poses(env, n, 0x5EED0005)
chk = capi.Checker(1)
t0 = time.time()
chk.set_environment(env)
chk.set_agent([agent])
print("scene upload %.2fs" % (time.time() - t0))
for opt in sys.argv[2:]:
    k, v = opt.split("=")
    chk.set_option(k, int(v))
chk.check_poses(poses)
t0 = time.perf_counter()
chk.check_poses(poses)
print("plain: %.3f ms host wall, kernel %.3f ms" % (1e3 * (time.perf_counter() - t0), chk.last_kernel_ms()))
chk.enable_counters(True)
chk.reset_counters()
bits = chk.check_poses(poses)
c = chk.counters()
print("gpu counters per pose:", {k: round(v / n, 3) for k, v in c.items()})
done, t_used, want, octr = bench.cpu_reference_run(env, agent, poses, os.cpu_count(), 20.0)
print("oracle per pose:", {k: round(v / done, 3) for k, v in octr.items()}, "cpu poses/s %.0f" % (done / t_used))
got = ((np.repeat(bits, 32) >> np.tile(np.arange(32, dtype=np.uint32), len(bits))) & 1).astype(bool)
print("mismatches:", int((got[:done] != want).sum()), "of", done, "collide frac", want.mean())
