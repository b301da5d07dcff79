#!/usr/bin/env python3
"""A small run of every split-pipeline variant for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from motionplanningtoolkit_b200 import capi, meshio, scenes  # noqa: E402

env = meshio.environment_submeshes(meshio.load_mptm("divider"))
agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
rng = scenes.SplitMix64(99)
poses = scenes.random_poses_near_objects(rng, 6000, env, 2.0, fmt="tq7")
ref = None
for opts in (dict(), dict(l2_persist=0), dict(qnode=0), dict(qnode=0, agent_top=255), dict(grid=0), dict(np2=1), dict(np3=0), dict(split=0)):
    chk = capi.Checker(1)
    chk.set_option("split_min", 32)
    for k, v in opts.items():
        chk.set_option(k, v)
    chk.set_environment(env)
    chk.set_agent([agent])
    bits = chk.check_poses(poses, capi.POSE_TQ7)
    if ref is None:
        ref = bits
    print(opts, "equal" if np.array_equal(bits, ref) else "DIFFERENT", int(np.unpackbits(bits.view(np.uint8)).sum()), flush=True)
    chk.close()
