#!/usr/bin/env python3
"""bench.py -- collision-checked poses/s on the BASELINE.json synthetic sweep (config C5).

  python bench.py --gpus N --steps K --warmup W            # this repo: libmptcc on N B200s (1 rank / GPU)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (FCL restated: oracle/)

A "step" is one pass of the hot path (MeshHandler::isInCollision per pose-set) over one batch of
2^20 synthetic pose-sets per GPU: blimp agent (all sub-meshes, 1 355 triangles) against a procedural
environment of 2 000 objects / ~1.14 M triangles.  Prints ONE JSON line (rank 0).
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
N_BATCHES = 4           # distinct pose batches rotated between steps (4 x 100.7 MB > 126 MB L2)
VERIFY_SAMPLE = 1 << 15


def workload_config(n_gpus, per_gpu):
    return {
        "workload": "C5 synthetic sweep: blimp.3ds (merge, 1355 tris) vs procedural environment "
                    "(%d objects, ~1.14M tris), %d pose-sets per GPU per step, 60%% uniform in the "
                    "environment bbox + 40%% near obstacles, uniform quaternions" % (N_OBJECTS, per_gpu),
        "pose_sets_per_step": per_gpu * n_gpus,
        "links": 1,
        "pose_format": "RT12 (fp64 3x4)",
        "l2": "inputs larger than L2: %d distinct pose batches (%.0f MB) rotate between steps"
              % (N_BATCHES, N_BATCHES * per_gpu * 96 / 1e6),
        "sharding": "pose-sets split contiguously per GPU, scene replicated, no collective",
    }


def make_scene():
    env = scenes.synthetic_environment(N_OBJECTS)
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    return env, agent


def make_poses(env, n, seed):
    rng = scenes.SplitMix64(seed)
    allv_lo = np.min([m[0].min(0) for m in env], axis=0)
    allv_hi = np.max([m[0].max(0) for m in env], axis=0)
    n_near = (n * 2) // 5
    p = np.concatenate([scenes.random_poses_rt12(rng, n - n_near, allv_lo, allv_hi),
                        scenes.random_poses_near_objects(rng, n_near, env, 3.0)])
    perm = np.argsort(rng.next_u64(n), kind="stable")
    return np.ascontiguousarray(p[perm])


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


def cpu_reference_run(env, agent, poses, threads, budget_s):
    """Time the oracle (FCL path restated; one isInCollision per pose-set, per-call object and broadphase
    tree construction included) on consecutive slices of `poses` until `budget_s` is used up."""
    from oracle import oracle as O
    models = {}
    om = [models.setdefault(id(m), O.Model(*m)) for m in env]
    oenv = O.Environment(om)
    olink = O.Model(*agent)
    ctr = O.Counters()
    done, t_used, bits = 0, 0.0, []
    chunk = 4096 * max(1, threads)
    while done < len(poses) and t_used < budget_s:
        sl = poses[done:done + chunk]
        t0 = time.perf_counter()
        b = O.verdicts(oenv, [olink], sl, O.POSE_RT12, threads, ctr)
        t_used += time.perf_counter() - t0
        bits.append(O.unpack_bits(b, len(sl)))
        done += len(sl)
    return done, t_used, np.concatenate(bits), ctr.as_dict()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    env, agent = make_scene()
    sample = 1 << 17
    poses = make_poses(env, sample, scenes.SEED_C5)
    times = []
    for i in range(args.warmup + args.steps):
        done, t_used, _, _ = cpu_reference_run(env, agent, poses, threads, 1e9)
        if i >= args.warmup:
            times.append(t_used)
    T = float(np.sum(times))
    value = sample * args.steps / T
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus, POSES_PER_GPU),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d pose-sets of the C5 batch per step, all %d host threads" % (sample, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference FCL 0.3.2 cannot be built here (not vendored); this is oracle/ (its fp64 restatement)",
    }
    print(json.dumps(line))
    return 0


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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

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
        p = make_poses(env, per_gpu, scenes.SEED_C5 + 7919 * rank + b)
        hp = capi.PinnedArray((per_gpu, 12), np.float64)
        hp.array[...] = p
        host_poses.append(hp)
        host_bits.append(capi.PinnedArray((n_words,), np.uint32))
        dev_poses.append(torch.from_numpy(p).cuda())
    dev_bits = torch.zeros(n_words, dtype=torch.int32, device="cuda")
    # a non-default stream: its handle is non-NULL (NULL would mean "the context's own stream" to the C ABI)
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()

    def step_device(i):
        chk.check_poses_device(per_gpu, dev_poses[i % N_BATCHES].data_ptr(), dev_bits.data_ptr(), capi.POSE_RT12, 0,
                               stream.cuda_stream)

    def step_host(i):
        b = i % N_BATCHES
        chk.check_poses(host_poses[b].array, capi.POSE_RT12, out=host_bits[b].array)

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
            t = chk.submit_poses(host_poses[b].array, host_bits[b].array, capi.POSE_RT12)
            if prev is not None:
                chk.wait(prev)
            prev = t
        chk.wait(prev)
        pipe_s = time.perf_counter() - t0
        barrier()
        pipe_s = max_over_ranks(pipe_s)
    clocks = clk.summary()
    # per-kernel device time of the pipeline (CUDA events inside the library, on the launching stream)
    phase_ms = None
    try:
        chk.set_option("time_phases", 1)
        acc = np.zeros(3)
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

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 traversal + f64 exact re-decide", "data": "synthetic",
        "config": dict(workload_config(world, per_gpu), host_cpu_affinity="gpu-local (NVML)" if numa_bound else "unbound"),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": per_gpu * 96 * world,
                "d2h_bytes_per_step": n_words * 4 * world, "ms_per_step": 1e3 * e2e_s / args.steps},
        "e2e_pipelined": {"value": total_poses / pipe_s, "unit": UNIT, "calls_in_flight": 2,
                          "ms_per_step": 1e3 * pipe_s / args.steps,
                          "note": "same host buffers and bytes as e2e through mptcc_submit_poses/mptcc_wait"},
        "gpu_launches": args.steps * (3 if per_gpu >= (1 << 15) else 1),
        "clocks": clocks,
    }

    if rank == 0:
        # ---- parity on a sample + CPU baseline + roofline (rank 0, N = 1 only for the CPU timing) -----
        os.sched_setaffinity(0, cpus_before)          # the CPU oracle may use every host core again
        threads = os.cpu_count() or 1
        sample_poses = host_poses[0].array            # whole batch 0 if the budget allows (about 2 s at 0.5 M poses/s)
        budget = 15.0 if world == 1 else 4.0
        done, t_used, want, octr = cpu_reference_run(env, agent, np.array(sample_poses), threads, budget)
        pipelined_words0 = np.array(host_bits[0].array)   # what the two-in-flight loop left for batch 0
        got_words = chk.check_poses(host_poses[0].array)
        got = ((np.repeat(got_words, 32) >> np.tile(np.arange(32, dtype=np.uint32), len(got_words))) & 1).astype(bool)
        mism = int((got[:done] != want).sum())
        line["parity"] = {"checked_pose_sets": int(done), "mismatches_vs_oracle": mism,
                          "collision_fraction": float(want.mean()),
                          "pipelined_equals_sync": bool(np.array_equal(pipelined_words0, got_words))}
        if world == 1:
            line["cpu_baseline"] = {"value": done / t_used, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "first %d pose-sets of batch 0 (of %d), oracle/ (FCL path restated in C, "
                                              "-O2 -ffp-contract=off), %d pthreads" % (done, per_gpu, threads)}
        # algorithmic bytes per pose-set, FCL-order test counts from the oracle sample
        bv = octr["n_bv_tests"] / max(1, done)
        lf = octr["n_leaf_tests"] / max(1, done)
        alg_bytes_per_pose = 96.0 + 0.125 + 64.0 * bv + 72.0 * lf
        step_ms = dev_ms / args.steps
        if phase_ms is not None:
            # dominant kernel = narrowphase_kernel: it does the OBB and triangle tests (64 n_bv + 72 n_leaf); the
            # 96-byte pose records are read by the broadphase kernel
            kernel_name, kernel_ms = "narrowphase_kernel<false, 128>", float(phase_ms[1])
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
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        line["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": traffic, "peak_source": peak_src,
                            "algorithmic_bytes_per_pose_set": alg_bytes_per_pose,
                            "fcl_order_bv_tests_per_pose_set": bv, "fcl_order_leaf_tests_per_pose_set": lf,
                            "kernel": kernel_name, "kernel_ms_per_launch": kernel_ms,
                            "algorithmic_bytes_per_pose_set_kernel": kernel_alg,
                            "pipeline_ms": {"broadphase": phase_ms[0], "narrowphase": phase_ms[1], "fallback": phase_ms[2]}
                            if phase_ms is not None else None,
                            "step_frac": alg_bytes_per_pose * per_gpu / (step_ms * 1e-3) / 1e9 / peak}
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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
