#!/usr/bin/env python3
"""bench.py -- collision-checked poses/s on the BASELINE.json synthetic sweep (config C5).

  python bench.py --gpus N --steps K --warmup W            # this repo: libmptcc on N B200s (1 rank / GPU)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (FCL restated: oracle/)
  python bench.py --config C1|C2|C3-last|C3-merge|C4|C5|C5-rt12|C5-terrain   # one BASELINE config at its full size,
                                                                             # whole-batch parity against the oracle

A "step" is one pass of the hot path (MeshHandler::isInCollision per pose-set) over one batch of
2^20 synthetic pose-sets per GPU: blimp agent (all sub-meshes, 1 355 triangles) against a procedural
environment of 2 000 objects / ~1.14 M triangles.  Poses are fcl::Transform3f(Quaternion3f, Vec3f) records
(TQ7, 56 bytes: what PRMLite builds for such samples, /root/reference/discretizations/workspace/prmlite.hpp:246-258);
the same batch as 3x4 matrices (RT12, 96 bytes) is reported next to it.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from motionplanningtoolkit_b200 import meshio, scenes  # noqa: E402

METRIC = "collision-checked poses/sec"
UNIT = "poses/s"
N_OBJECTS = 2000
POSES_PER_GPU = 1 << 20
N_BATCHES = 4           # distinct pose batches rotated between steps (4 x 58.7 MB plus 200 MB of tree nodes > 126 MB L2)
FMT_NAMES = {"rt12": 0, "tq7": 1, "t3": 2, "xyzcs": 3}
REC_LEN = {0: 12, 1: 7, 2: 3, 3: 5}


def workload_config(n_gpus, per_gpu, fmt="tq7"):
    return {
        "workload": "C5 synthetic sweep: blimp.3ds (merge, 1355 tris) vs procedural environment "
                    "(%d objects, ~1.14M tris), %d pose-sets per GPU per step, 60%% uniform in the "
                    "environment bbox + 40%% near obstacles, uniform quaternions" % (N_OBJECTS, per_gpu),
        "pose_sets_per_step": per_gpu * n_gpus,
        "links": 1,
        "pose_format": {"tq7": "TQ7 (fp64 x y z qw qx qy qz = fcl::Transform3f(q, T), 56 B)",
                        "rt12": "RT12 (fp64 3x4 = fcl::Transform3f(R, T), 96 B)"}[fmt],
        "l2": "inputs larger than L2: %d distinct pose batches (%.0f MB) rotate between steps, next to 200 MB of hot tree data"
              % (N_BATCHES, N_BATCHES * per_gpu * 8 * REC_LEN[FMT_NAMES[fmt]] / 1e6),
        "sharding": "pose-sets split contiguously per GPU, scene replicated, no collective",
    }


def make_scene():
    env = scenes.synthetic_environment(N_OBJECTS)
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    return env, agent


def make_poses(env, n, seed, fmt="rt12"):
    """The C5 batch.  The same seed gives the same translations and quaternions in either record format."""
    rng = scenes.SplitMix64(seed)
    allv_lo = np.min([m[0].min(0) for m in env], axis=0)
    allv_hi = np.max([m[0].max(0) for m in env], axis=0)
    n_near = (n * 2) // 5
    a = scenes.random_poses_tq7(rng, n - n_near, allv_lo, allv_hi)
    b = scenes.random_poses_near_objects(rng, n_near, env, 3.0, fmt="tq7")
    p = np.concatenate([a, b])
    perm = np.argsort(rng.next_u64(n), kind="stable")
    p = p[perm]
    if fmt == "rt12":
        p = scenes.poses_rt12(scenes.quat_to_matrix(p[:, 3:7]), p[:, :3])
    return np.ascontiguousarray(p)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples = []
        self.stop = threading.Event()
        self.th = None

    def _run(self):
        nv = None
        try:                                   # NVML is much faster to poll than spawning nvidia-smi
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            nv = (pynvml, h)
        except Exception:
            nv = None
        while not self.stop.is_set():
            try:
                if nv is not None:
                    m, h = nv
                    sm = m.nvmlDeviceGetClockInfo(h, m.NVML_CLOCK_SM)
                    mx = m.nvmlDeviceGetMaxClockInfo(h, m.NVML_CLOCK_SM)
                    pw = m.nvmlDeviceGetPowerUsage(h) / 1000.0
                    r = m.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    act = lambda bit: "Active" if (r & bit) else "Not Active"
                    self.samples.append([str(sm), str(mx), str(pw), act(m.nvmlClocksThrottleReasonHwSlowdown),
                                         act(m.nvmlClocksThrottleReasonHwThermalSlowdown),
                                         act(m.nvmlClocksThrottleReasonSwThermalSlowdown),
                                         act(m.nvmlClocksThrottleReasonSwPowerCap)])
                    self.stop.wait(0.005)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[3 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


def _to_oracle_format(poses, fmt):
    """The oracle speaks RT12 and TQ7; the compact formats are by definition the RT12 records with the same entries."""
    if fmt in (0, 1):
        return poses, fmt
    p = np.asarray(poses, dtype=np.float64)
    lead = p.shape[:-1]
    p = p.reshape(-1, REC_LEN[fmt])
    if fmt == 2:
        out = scenes.translation_poses(p)
    else:
        R = np.tile(np.eye(3), (len(p), 1, 1))
        R[:, 0, 0], R[:, 0, 1], R[:, 1, 0], R[:, 1, 1] = p[:, 3], p[:, 4], -p[:, 4], p[:, 3]
        out = scenes.poses_rt12(R, p[:, :3])
    return out.reshape(lead + (12,)), 0


def cpu_reference_run(env, links, poses, threads, budget_s, fmt=0, pose7=(0, 0, 0, 1, 0, 0, 0)):
    """Time the oracle (FCL path restated; one isInCollision per pose-set, per-call object and broadphase
    tree construction included) on consecutive slices of `poses` until `budget_s` is used up."""
    from oracle import oracle as O
    if not isinstance(links, list):
        links = [links]
    models = {}
    om = [models.setdefault(id(m), O.Model(*m)) for m in env]
    oenv = O.Environment(om, pose7)
    olinks = [models.setdefault(id(m), O.Model(*m)) for m in links]
    poses, ofmt = _to_oracle_format(poses, fmt)
    ctr = O.Counters()
    done, t_used, bits = 0, 0.0, []
    chunk = 4096 * max(1, threads)
    while done < len(poses) and t_used < budget_s:
        sl = poses[done:done + chunk]
        t0 = time.perf_counter()
        b = O.verdicts(oenv, olinks, sl, ofmt, threads, ctr)
        t_used += time.perf_counter() - t0
        bits.append(O.unpack_bits(b, len(sl)))
        done += len(sl)
    return done, t_used, np.concatenate(bits), ctr.as_dict()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    fmt = args.pose_format
    env, agent = make_scene()
    sample = 1 << 17
    poses = make_poses(env, sample, scenes.SEED_C5, fmt)
    times = []
    for i in range(args.warmup + args.steps):
        done, t_used, _, _ = cpu_reference_run(env, agent, poses, threads, 1e9, FMT_NAMES[fmt])
        if i >= args.warmup:
            times.append(t_used)
    T = float(np.sum(times))
    value = sample * args.steps / T
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus, args.poses_per_gpu, fmt),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d pose-sets of the C5 batch per step, all %d host threads" % (sample, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference FCL 0.3.2 cannot be built here (not vendored); this is oracle/ (its fp64 restatement)",
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------
# the other BASELINE configs at their full sizes
# ---------------------------------------------------------------------------------------------------------------
def build_config(name, n_override=None):
    """-> dict(env, links, poses [n, L, rec], fmt, pose7, what)"""
    box = meshio.load_mptm("unit_box")[0]
    ident = (0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0)
    if name == "C1":
        n = n_override or (1 << 20)
        rng = scenes.SplitMix64(scenes.SEED_C1)
        n_edges = n // 10
        half = n_edges // 2                         # uniform starts in the bounds + starts within 3 of the obstacle
        S = np.concatenate([np.stack([rng.uniform(half, -10, 10) for _ in range(3)], 1),
                            np.stack([rng.uniform(n_edges - half, -3, 3) for _ in range(3)], 1)])
        D = np.stack([rng.uniform(n_edges, -1, 1) for _ in range(3)], 1)
        D /= np.sqrt((D * D).sum(1))[:, None]
        pts, _ = scenes.omnidirectional_edges_poses(S, S + D, 0.1)
        return dict(env=[box], links=[box], poses=pts[:, None, :], fmt=2, pose7=ident,
                    what="omnidirectional.inst: unit_box vs unit_box, RRT unit steps at dt 0.1 (T3 records)")
    if name == "C2":
        rng = scenes.SplitMix64(scenes.SEED_C2)
        pts_all, chunks = None, []
        for _ in range(8):
            pts = np.stack([rng.uniform(100, -10, 10) for _ in range(3)], axis=1)
            allp = pts if pts_all is None else np.concatenate([pts_all, pts])
            first = 0 if pts_all is None else len(pts_all)
            S, E = [], []
            for i in range(max(first, 1), len(allp)):
                d = np.linalg.norm(allp[:i] - allp[i], axis=1)
                for j in np.argsort(d, kind="stable")[:10]:
                    S.append(allp[i])
                    E.append(allp[j])
            p, _ = scenes.omnidirectional_edges_poses(np.array(S), np.array(E), 0.1)
            chunks.append(p)
            pts_all = allp
        pts = np.concatenate(chunks)
        return dict(env=[box], links=[box], poses=pts[:, None, :], fmt=2, pose7=ident,
                    what="PRM roadmap: 8 slices x 100 milestones, k=10 straight edges at dt 0.1 (T3 records)")
    if name in ("C3-last", "C3-merge"):
        n = n_override or (1 << 20)
        env = meshio.environment_submeshes(meshio.load_mptm("divider"))
        agent = meshio.agent_submesh(meshio.load_mptm("blimp"), name[3:])
        lo = np.min([m[0].min(0) for m in env], axis=0)
        hi = np.max([m[0].max(0) for m in env], axis=0)
        rng = scenes.SplitMix64(scenes.SEED_C3)
        p12, _ = scenes.blimp_edge_poses(rng, n // 10, lo, hi)
        # the same records as x y z cos sin (the entries of stateToFCLTransform's matrix)
        xyzcs = np.stack([p12[:, 9], p12[:, 10], p12[:, 11], p12[:, 0], p12[:, 1]], axis=1)
        return dict(env=env, links=[agent], poses=xyzcs[:, None, :], fmt=3, pose7=ident,
                    what="blimp.inst: blimp.3ds (%s) vs divider.dae (110 objects), doStep edges, 10 poses each (XYZCS records)" % name[3:])
    if name == "C4":
        n = n_override or (1 << 18)
        rng = scenes.SplitMix64(scenes.SEED_C4)
        p12 = scenes.snake_state_poses(scenes.snake_random_states(rng, n, (-4, -4), (4, 4)))
        xyzcs = np.stack([p12[..., 9], p12[..., 10], p12[..., 11], p12[..., 0], p12[..., 1]], axis=-1)
        return dict(env=[box], links=[box] * 11, poses=xyzcs, fmt=3, pose7=ident,
                    what="snake.inst: 11 unit_box links per pose-set (stateToFCLTransforms verbatim), obstacle at the origin (XYZCS records)")
    if name in ("C5", "C5-rt12"):
        n = n_override or (1 << 20)
        env, agent = make_scene()
        fmt = "rt12" if name.endswith("rt12") else "tq7"
        return dict(env=env, links=[agent], poses=make_poses(env, n, scenes.SEED_C5, fmt)[:, None, :], fmt=FMT_NAMES[fmt],
                    pose7=ident, what="C5 synthetic sweep (%s records)" % fmt.upper())
    if name == "C5-terrain":
        n = n_override or (1 << 20)
        env = [scenes.terrain(708, 708)]
        agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
        rng = scenes.SplitMix64(scenes.SEED_C5 + 1)
        p = scenes.random_poses_tq7(rng, n, (-95, -95, -10), (95, 95, 12))
        return dict(env=env, links=[agent], poses=p[:, None, :], fmt=1, pose7=ident,
                    what="C5 single-object variant: blimp vs one 1.0 M-triangle terrain (TQ7 records)")
    raise SystemExit("unknown config %s" % name)


def run_config(name, capi, oracle_budget_s, n_override=None, reps=3):
    """One BASELINE config at its full size: pose-sets/s through mptcc_check_poses (pinned host buffers), device
    milliseconds, and parity against the oracle on as much of the batch as `oracle_budget_s` allows."""
    cfg = build_config(name, n_override)
    poses = np.ascontiguousarray(cfg["poses"], dtype=np.float64)
    n, fmt = poses.shape[0], cfg["fmt"]
    chk = capi.Checker(1)
    chk.set_environment(cfg["env"], cfg["pose7"])
    chk.set_agent(cfg["links"])
    pin = capi.PinnedArray(poses.shape, np.float64)
    pin.array[...] = poses
    out = capi.PinnedArray(((n + 31) // 32,), np.uint32)
    chk.check_poses(pin.array, fmt, out=out.array)
    t0 = time.perf_counter()
    for _ in range(reps):
        chk.check_poses(pin.array, fmt, out=out.array)
    dt = (time.perf_counter() - t0) / reps
    dev_ms = chk.last_kernel_ms()
    got = np.unpackbits(out.array.view(np.uint8), bitorder="little")[:n].astype(bool)
    threads = os.cpu_count() or 1
    done, t_used, want, _ = cpu_reference_run(cfg["env"], cfg["links"], poses, threads, oracle_budget_s, fmt, cfg["pose7"])
    res = {"what": cfg["what"], "pose_sets": int(n), "links": len(cfg["links"]), "bytes_per_pose_set": 8 * REC_LEN[fmt] * len(cfg["links"]),
           "pose_sets_per_s_host_buffers": n / dt, "device_ms": dev_ms, "collision_fraction": float(got.mean()),
           "checked_vs_oracle": int(done), "mismatches_vs_oracle": int((got[:done] != want).sum()),
           "cpu_oracle_pose_sets_per_s": done / t_used, "cpu_threads": threads}
    chk.close()
    pin.free()
    out.free()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    from motionplanningtoolkit_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)

    if args.config:                                 # one BASELINE config, whole-batch parity (rank 0 / 1 GPU)
        if rank == 0:
            res = run_config(args.config, capi, 1e9, None)
            print(json.dumps({"config": args.config, **res}))
        return 0

    # N > 1: run this rank (and first-touch its pinned staging buffers) on the CPUs next to its GPU, so that
    # eight ranks do not pull their host buffers across the socket interconnect
    cpus_before = os.sched_getaffinity(0)
    numa_bound = False
    if world > 1:
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
            numa_bound = os.sched_getaffinity(0) != cpus_before
        except Exception:
            numa_bound = False
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # a CPU-side group: a rank that waits in an NCCL barrier keeps a spinning kernel on its GPU, which must not
        # happen while rank 0 drives ALL the GPUs from one context (the e2e_one_context leg below)
        cpu_group = dist.new_group(backend="gloo")

    def cpu_barrier():
        if world > 1:
            dist.barrier(group=cpu_group)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    per_gpu = args.poses_per_gpu
    fmt_name = args.pose_format
    fmt, rec = FMT_NAMES[fmt_name], REC_LEN[FMT_NAMES[fmt_name]]
    alt_name = "rt12" if fmt_name == "tq7" else "tq7"
    afmt, arec = FMT_NAMES[alt_name], REC_LEN[FMT_NAMES[alt_name]]
    env, agent = make_scene()
    chk = capi.Checker(device_ids=[local])
    for kv in filter(None, os.environ.get("MPTCC_OPTS", "").split(",")):   # tuning experiments only
        k, v = kv.split("=")
        chk.set_option(k, int(v))
    chk.set_environment(env)
    chk.set_agent([agent])

    n_words = (per_gpu + 31) // 32
    host_poses, host_bits, dev_poses = [], [], []
    for b in range(N_BATCHES):
        p = make_poses(env, per_gpu, scenes.SEED_C5 + 7919 * rank + b, fmt_name)
        hp = capi.PinnedArray((per_gpu, rec), np.float64)
        hp.array[...] = p
        host_poses.append(hp)
        host_bits.append(capi.PinnedArray((n_words,), np.uint32))
        dev_poses.append(torch.from_numpy(p).cuda())
    alt_host = capi.PinnedArray((per_gpu, arec), np.float64)       # batch 0 in the other record format
    alt_host.array[...] = make_poses(env, per_gpu, scenes.SEED_C5 + 7919 * rank, alt_name)
    alt_bits = capi.PinnedArray((n_words,), np.uint32)
    dev_bits = torch.zeros(n_words, dtype=torch.int32, device="cuda")
    # a non-default stream: its handle is non-NULL (NULL would mean "the context's own stream" to the C ABI)
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()

    def step_device(i):
        chk.check_poses_device(per_gpu, dev_poses[i % N_BATCHES].data_ptr(), dev_bits.data_ptr(), fmt, 0, stream.cuda_stream)

    def step_host(i):
        b = i % N_BATCHES
        chk.check_poses(host_poses[b].array, fmt, out=host_bits[b].array)

    # ---- kernel-only: inputs resident in HBM --------------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
    torch.cuda.synchronize()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        torch.cuda.synchronize()
        ev0.record(stream)
        for i in range(args.steps):
            step_device(i)
        ev1.record(stream)
        torch.cuda.synchronize()
        barrier()
        dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
        # ---- end to end: pinned host poses in, verdict words back on the host, every step ------------
        for i in range(max(1, min(args.warmup, 3))):
            step_host(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step_host(i)
        e2e_s = time.perf_counter() - t0
        barrier()
        e2e_s = max_over_ranks(e2e_s)
        # ---- the same host-buffer calls with two in flight (mptcc_submit_poses / mptcc_wait): the copy of
        #      step i+1 overlaps the kernels of step i.  Reported next to `e2e`, not instead of it. ---------
        t0 = time.perf_counter()
        prev = None
        for i in range(args.steps):
            b = i % N_BATCHES
            t = chk.submit_poses(host_poses[b].array, host_bits[b].array, fmt)
            if prev is not None:
                chk.wait(prev)
            prev = t
        chk.wait(prev)
        pipe_s = time.perf_counter() - t0
        barrier()
        pipe_s = max_over_ranks(pipe_s)
        # ---- the other record format of the same poses (host buffers, blocking calls) ------------------
        alt_steps = max(3, min(args.steps, 20))
        chk.check_poses(alt_host.array, afmt, out=alt_bits.array)
        barrier()
        t0 = time.perf_counter()
        for i in range(alt_steps):
            chk.check_poses(alt_host.array, afmt, out=alt_bits.array)
        alt_s = time.perf_counter() - t0
        barrier()
        alt_s = max_over_ranks(alt_s)
    clocks = clk.summary()
    # per-kernel device time of the pipeline (CUDA events inside the library, on the launching stream)
    phase_ms = None
    try:
        chk.set_option("time_phases", 1)
        acc = np.zeros(4)
        n_ph = max(3, min(args.steps, 10))
        for i in range(n_ph):
            step_device(i)
            acc += np.array(chk.last_phase_ms(0))
        phase_ms = (acc / n_ph).tolist()
        chk.set_option("time_phases", 0)
    except Exception:
        phase_ms = None
    total_poses = per_gpu * world * args.steps
    value = total_poses / (dev_ms * 1e-3)
    e2e_value = total_poses / e2e_s

    _opts_env = os.environ.get("MPTCC_OPTS", "") + "," + os.environ.get("MPTCC_OPTIONS", "")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 traversal + f64 exact re-decide", "data": "synthetic",
        "config": workload_config(world, per_gpu, fmt_name),
        "host_cpu_affinity": "gpu-local (NVML)" if numa_bound else "unbound",
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": per_gpu * 8 * rec * world,
                "d2h_bytes_per_step": n_words * 4 * world, "ms_per_step": 1e3 * e2e_s / args.steps},
        "e2e_pipelined": {"value": total_poses / pipe_s, "unit": UNIT, "calls_in_flight": 2,
                          "ms_per_step": 1e3 * pipe_s / args.steps,
                          "note": "same host buffers and bytes as e2e through mptcc_submit_poses/mptcc_wait"},
        "e2e_" + alt_name: {"value": per_gpu * world * alt_steps / alt_s, "unit": UNIT, "h2d_bytes_per_step": per_gpu * 8 * arec * world,
                            "ms_per_step": 1e3 * alt_s / alt_steps, "note": "the same poses as %s records, blocking calls" % alt_name.upper()},
        # kernels of this library launched inside the device-timed region, all ranks: a split-pipeline step is broadphase +
        # narrowphase + exact pass + (early-exit) fall-back with the np3 / np2 narrowphase, one launch fewer with the round-1
        # kernel (its exact pass is inline); a batch below the split threshold is one fused launch
        "gpu_launches": args.steps * world * ((3 if ("np3=0" in _opts_env and "np2=1" not in _opts_env) else 4) if per_gpu >= (1 << 15) else 1),
        "clocks": clocks,
    }

    # ---- one process, N devices: the library's own multi-GPU context (mptcc_create(n_devices = N)) -------------
    words_rank0 = []
    if world > 1:
        if rank == 0:
            for b in range(N_BATCHES):
                words_rank0.append(np.array(chk.check_poses(host_poses[b].array, fmt)))
        else:
            chk.close()                             # free this rank's device memory and staging buffers
        torch.cuda.synchronize()
        cpu_barrier()
        if rank == 0:
            big = capi.PinnedArray((per_gpu * world, rec), np.float64)
            for r in range(world):
                big.array[r * per_gpu:(r + 1) * per_gpu] = host_poses[r % N_BATCHES].array
            big_bits = capi.PinnedArray((n_words * world,), np.uint32)
            one = capi.Checker(device_ids=list(range(world)))
            one.set_environment(env)
            one.set_agent([agent])
            one.check_poses(big.array, fmt, out=big_bits.array)
            oc_steps = max(3, min(args.steps, 10))
            t0 = time.perf_counter()
            for _ in range(oc_steps):
                one.check_poses(big.array, fmt, out=big_bits.array)
            oc_s = time.perf_counter() - t0
            same = all(np.array_equal(big_bits.array[r * n_words:(r + 1) * n_words], words_rank0[r % N_BATCHES]) for r in range(world))
            line["e2e_one_context"] = {"value": per_gpu * world * oc_steps / oc_s, "unit": UNIT, "devices": world,
                                       "pose_sets_per_call": per_gpu * world, "ms_per_call": 1e3 * oc_s / oc_steps,
                                       "equals_per_rank_words": bool(same),
                                       "note": "ONE process, ONE mptcc context over all %d GPUs (mptcc_create(n_devices)), "
                                               "one blocking mptcc_check_poses per step on pinned host buffers" % world}
            one.close()
        cpu_barrier()

    if rank == 0:
        # ---- parity on the whole batch + CPU baseline + roofline (rank 0, N = 1 only for the CPU timing) -----
        os.sched_setaffinity(0, cpus_before)          # the CPU oracle may use every host core again
        threads = os.cpu_count() or 1
        budget = 15.0 if world == 1 else 4.0
        done, t_used, want, octr = cpu_reference_run(env, agent, np.array(host_poses[0].array), threads, budget, fmt)
        pipelined_words0 = np.array(host_bits[0].array)   # what the two-in-flight loop left for batch 0
        got_words = chk.check_poses(host_poses[0].array, fmt) if world == 1 else words_rank0[0]
        unpack = lambda w: np.unpackbits(np.ascontiguousarray(w).view(np.uint8), bitorder="little").astype(bool)
        got = unpack(got_words)
        mism = int((got[:done] != want).sum())
        line["parity"] = {"checked_pose_sets": int(done), "mismatches_vs_oracle": mism,
                          "collision_fraction": float(want.mean()),
                          "pipelined_equals_sync": bool(np.array_equal(pipelined_words0, got_words))}
        if world == 1:
            # the TIMED device-resident path: its verdict words for batch 0 against the same oracle bits
            dev_bits.zero_()
            step_device(0)
            torch.cuda.synchronize()
            dgot = unpack(dev_bits.cpu().numpy().view(np.uint32))
            line["parity"]["device_path_mismatches_vs_oracle"] = int((dgot[:done] != want).sum())
            # the other record format has its own oracle run (R is derived from q / q from R)
            adone, _, awant, _ = cpu_reference_run(env, agent, np.array(alt_host.array), threads, 4.0, afmt)
            agot = unpack(alt_bits.array)
            line["parity"][alt_name + "_checked"] = int(adone)
            line["parity"][alt_name + "_mismatches_vs_oracle"] = int((agot[:adone] != awant).sum())
            mism += line["parity"]["device_path_mismatches_vs_oracle"] + line["parity"][alt_name + "_mismatches_vs_oracle"]
            d1, t1, _, _ = cpu_reference_run(env, agent, np.array(host_poses[0].array[:1 << 14]), 1, 4.0, fmt)
            line["cpu_baseline"] = {"value": done / t_used, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "first %d pose-sets of batch 0 (of %d), oracle/ (FCL path restated in C, "
                                              "-O2 -ffp-contract=off), %d pthreads" % (done, per_gpu, threads),
                                    "cpu_ref_1t": {"value": d1 / t1, "unit": UNIT, "cores": 1,
                                                   "sample": "first %d pose-sets of batch 0, one thread (the toolkit itself is "
                                                             "single-threaded)" % d1}}
        # algorithmic bytes per pose-set, FCL-order test counts from the oracle sample
        bv = octr["n_bv_tests"] / max(1, done)
        lf = octr["n_leaf_tests"] / max(1, done)
        in_bytes = 8.0 * rec
        alg_bytes_per_pose = in_bytes + 0.125 + 64.0 * bv + 72.0 * lf
        alg_flops_per_pose = 330.0 * bv + 800.0 * lf
        step_ms = dev_ms / args.steps
        opts_env = os.environ.get("MPTCC_OPTS", "") + "," + os.environ.get("MPTCC_OPTIONS", "")
        np2 = "np2=1" in opts_env
        np1 = "np3=0" in opts_env and not np2
        if phase_ms is not None:
            # dominant kernel = the narrowphase kernel: it does the OBB and triangle tests (64 n_bv + 72 n_leaf); the
            # pose records are read by the broadphase kernel
            kernel_name = ("narrowphase2_kernel<false, 128, false>" if np2 else "narrowphase_kernel<false, 128>" if np1
                           else "narrowphase3_kernel<false, false, true>")
            kernel_ms = float(phase_ms[1])
            kernel_alg = 64.0 * bv + 72.0 * lf
        else:
            kernel_name, kernel_ms, kernel_alg = "collide_kernel<false>", step_ms, alg_bytes_per_pose
        achieved = kernel_alg * per_gpu / (kernel_ms * 1e-3) / 1e9
        peak, peak_src = 6650.0, "fallback"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured"
        except Exception:
            pass
        pk = {}
        try:
            with open(os.path.join(ROOT, "profiles", "r2_peaks.json")) as f:
                pk = json.load(f)
        except Exception:
            pk = {}
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        tflops = alg_flops_per_pose * per_gpu / (kernel_ms * 1e-3) / 1e12
        fracs = {"hbm": achieved / peak}
        if pk.get("l2_read_gbs"):
            fracs["l2"] = achieved / pk["l2_read_gbs"]
        if pk.get("fp32_tflops"):
            fracs["fp32"] = tflops / pk["fp32_tflops"]
        top = max(fracs, key=fracs.get)
        latency = max(fracs.values()) < 0.5
        # ncu figures quoted in the note come from the newest committed summary of this kernel (profiles/, tools/make_profiles.py)
        ncu_note = "no ncu summary of this kernel under profiles/"
        try:
            import glob
            mfile = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_narrowphase3_kernel_metrics.txt")))[-1]
            mv = {}
            for ln in open(mfile):
                t = ln.split()
                if len(t) >= 2:
                    mv[t[0]] = t[-1].replace(",", "")
            ncu_note = ("ncu (profiles/%s): %s registers, %.0f %% of the warp slots active, issue slots %.0f %% busy, fp32 pipe %.0f %%, "
                        "L1 data pipe %.0f %%, L2 hit rate %.0f %%, DRAM %.0f %% of peak" % (
                            os.path.basename(mfile), mv["launch__registers_per_thread"],
                            float(mv["sm__warps_active.avg.pct_of_peak_sustained_active"]),
                            float(mv["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                            float(mv["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]),
                            float(mv["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]),
                            float(mv["lts__t_sector_hit_rate.pct"]),
                            float(mv["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"])))
        except Exception:
            pass
        line["roofline"] = {"bound": "hbm",
                            "bound_measured": "latency" if latency else top,
                            "bound_note": ("the contract's HBM roofline is reported in `frac` (algorithmic bytes = FCL-order tests x "
                                           "64-byte fp32 nodes; the kernel itself gathers 32-byte quantised nodes, mostly from L1/L2: "
                                           "DRAM traffic is `traffic`); the kernel is bound by instruction issue and dependent latency "
                                           "at 16 warps per SM, not by a memory level; " + ncu_note),
                            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "hbm_frac": fracs["hbm"], "l2_frac": fracs.get("l2"), "fp32_frac": fracs.get("fp32"),
                            "l2_gather_frac": achieved / pk["gather64_l2_gbs"] if pk.get("gather64_l2_gbs") else None,
                            "achieved_fp32_tflops": tflops,
                            "peaks": {"hbm_gbs": peak, "hbm_source": peak_src, "l2_read_gbs": pk.get("l2_read_gbs"),
                                      "l2_gather64_gbs": pk.get("gather64_l2_gbs"), "fp32_tflops": pk.get("fp32_tflops"),
                                      "source": "profiles/r2_peaks.json (tools/peaks.cu, measured on this pool's B200)"},
                            "traffic": traffic, "peak_source": peak_src,
                            "algorithmic_bytes_per_pose_set": alg_bytes_per_pose,
                            "algorithmic_flops_per_pose_set": alg_flops_per_pose,
                            "fcl_order_bv_tests_per_pose_set": bv, "fcl_order_leaf_tests_per_pose_set": lf,
                            "kernel": kernel_name, "kernel_ms_per_launch": kernel_ms,
                            "algorithmic_bytes_per_pose_set_kernel": kernel_alg,
                            "pipeline_ms": {"broadphase": phase_ms[0], "narrowphase": phase_ms[1], "exact": phase_ms[2],
                                            "fallback": phase_ms[3]} if phase_ms is not None else None,
                            "step_frac": alg_bytes_per_pose * per_gpu / (step_ms * 1e-3) / 1e9 / peak}
        # ---- the other BASELINE configs at their full sizes (N = 1 only: a few seconds each) -----------------------
        if world == 1 and not args.no_other_configs:
            chk.close()
            others = {}
            for name in ("C1", "C2", "C3-last", "C3-merge", "C4"):
                try:
                    others[name] = run_config(name, capi, 3.0)
                    mism += others[name]["mismatches_vs_oracle"]
                except Exception as ex:                                  # never lose the headline line
                    others[name] = {"error": repr(ex)}
            line["other_configs"] = others
        if mism:
            line["invalid"] = "verdict mismatches vs oracle"
        print(json.dumps(line))
    barrier()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--poses-per-gpu", type=int, default=POSES_PER_GPU)
    ap.add_argument("--pose-format", default="tq7", choices=["tq7", "rt12"])
    ap.add_argument("--config", default=None, help="run ONE BASELINE config at full size with whole-batch parity")
    ap.add_argument("--no-other-configs", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
