#!/usr/bin/env python3
"""Time mptcc_check_poses on one of the fixture scenes (BASELINE configs C1/C3/C4 as perf cases).
usage: perf_scene.py <scene> [n_poses] [opt=value ...]   scene: c1 | c3last | c3merge | c4 | terrain"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from motionplanningtoolkit_b200 import capi, meshio, scenes  # noqa: E402

scene = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
rng = scenes.SplitMix64(0xBEEF)
box = meshio.load_mptm("unit_box")[0]
if scene == "c1":
    env, links = [box], [box]
    poses = scenes.translation_poses(np.stack([rng.uniform(n, -3, 3) for _ in range(3)], 1))
elif scene in ("c3last", "c3merge"):
    env = meshio.environment_submeshes(meshio.load_mptm("divider"))
    links = [meshio.agent_submesh(meshio.load_mptm("blimp"), "last" if scene == "c3last" else "merge")]
    lo = np.min([m[0].min(0) for m in env], axis=0)
    hi = np.max([m[0].max(0) for m in env], axis=0)
    poses, _ = scenes.blimp_edge_poses(rng, n // 10, lo, hi)
elif scene == "c4":
    env, links = [box], [box] * 11
    poses = scenes.snake_state_poses(scenes.snake_random_states(rng, n, (-4, -4), (4, 4)))
elif scene == "terrain":
    env = [scenes.terrain(708, 708)]
    links = [meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")]
    poses = scenes.random_poses_rt12(rng, n, (-95, -95, -10), (95, 95, 12))
else:
    raise SystemExit(__doc__)
n = len(poses)
chk = capi.Checker(1)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    chk.set_option(k, int(v))
chk.set_environment(env)
chk.set_agent(links)
pin = capi.PinnedArray(poses.shape, np.float64)
pin.array[...] = poses
out = capi.PinnedArray(((n + 31) // 32,), np.uint32)
for _ in range(3):
    chk.check_poses(pin.array, out=out.array)
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    chk.check_poses(pin.array, out=out.array)
dt = (time.perf_counter() - t0) / reps
hits = int(np.unpackbits(out.array.view(np.uint8)).sum())
chk.enable_counters(True)
chk.reset_counters()
chk.check_poses(pin.array, out=out.array)
c = chk.counters()
print("%-8s n=%d links=%d  e2e %.3f ms  %.1f Mposes/s  device %.3f ms  collide %.3f  cand/pose %.2f  bv/pose %.1f  leaf/pose %.2f  exact/pose %.3f"
      % (scene, n, len(links), dt * 1e3, n / dt / 1e6, chk.last_kernel_ms(), hits / n, c["n_candidates"] / n, c["n_bv_tests"] / n,
         c["n_leaf_tests"] / n, c["n_exact_tri"] / n))
if any(a.startswith("time_phases=1") for a in sys.argv[3:]):
    chk.enable_counters(False)
    chk.check_poses(pin.array, out=out.array)
    print("         phases of the last piece (broadphase, narrowphase, exact, fall-back) ms:", ["%.3f" % v for v in chk.last_phase_ms(0)])
