#!/usr/bin/env python3
"""A/B of narrowphase builds on the C5 workload: device-timed phases + a checksum of the verdict words.
usage: MPTCC_LIB=<lib> python tools/ab_np.py [n_poses] [opt=value ...]"""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from motionplanningtoolkit_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1 << 20
opts = [a for a in sys.argv[1:] if "=" in a]
env, agent = bench.make_scene()
chk = capi.Checker(1)
chk.set_environment(env)
chk.set_agent([agent])
for kv in opts:
    k, v = kv.split("=")
    chk.set_option(k, int(v))
batches = [torch.from_numpy(bench.make_poses(env, n, 0x5EED0005 + b)).cuda() for b in range(3)]
bits = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
stream = torch.cuda.Stream()
chk.set_option("time_phases", 1)
acc = np.zeros(4)
reps = 12
for i in range(3 + reps):
    chk.check_poses_device(n, batches[i % 3].data_ptr(), bits.data_ptr(), capi.POSE_RT12, 0, stream.cuda_stream)
    ph = np.array(chk.last_phase_ms(0))
    if i >= 3:
        acc += ph
torch.cuda.synchronize()
chk.check_poses_device(n, batches[0].data_ptr(), bits.data_ptr(), capi.POSE_RT12, 0, stream.cuda_stream)
torch.cuda.synchronize()
h = hashlib.sha256(bits.cpu().numpy().tobytes()).hexdigest()[:12]
acc /= reps
# one more launch with the counters on: OBB / leaf tests per pose-set (not timed)
chk.enable_counters(True)
chk.reset_counters()
chk.check_poses_device(n, batches[0].data_ptr(), bits.data_ptr(), capi.POSE_RT12, 0, stream.cuda_stream)
torch.cuda.synchronize()
cd = chk.counters()
h2 = hashlib.sha256(bits.cpu().numpy().tobytes()).hexdigest()[:12]
tests = " bv/pose %.1f leaf/pose %.2f exact/pose %.3f%s" % (cd.get("n_bv_tests", 0) / n, cd.get("n_leaf_tests", 0) / n,
                                                          cd.get("n_exact_tri", 0) / n, "" if h2 == h else " COUNT-BUILD BITS DIFFER " + h2)
print("%-28s %s bp %.3f np %.3f exact %.3f fb %.3f total %.3f ms  bits %s" % (
    os.path.basename(os.environ.get("MPTCC_LIB", "libmptcc.so")), " ".join(opts), acc[0], acc[1], acc[2], acc[3], acc.sum(), h) + tests)
