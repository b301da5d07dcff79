#!/usr/bin/env python3
"""Summarise ncu captures (gpurun_out/) into profiles/ (tracked).  usage: make_profiles.py <round-tag>"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sass__inst_executed_local_loads',
        'sass__inst_executed_local_stores', 'sm__cycles_elapsed.avg.per_second',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def launches():
    src = os.path.join(G, "launches_%s.csv" % tag)
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ci = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    with open(os.path.join(P, "%s_launches.csv" % tag), "w") as f:
        for r in rows[hi:]:
            f.write(",".join('"%s"' % c for c in r) + "\n")
    for r in rows[hi + 1:]:
        if len(r) != len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ci["Metric Value"]].replace(",", ""))
        u = r[ci["Metric Unit"]]
        ms = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
        k = r[ci["Kernel Name"]][:100]
        agg[k][0] += 1
        agg[k][1] += ms
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, "%s_launch_summary.txt" % tag), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 200 ; command: python bench.py --steps 5 --warmup 3 --no-other-configs\n")
        f.write("(cold-cache, serialised launch times: compare SHARES, not absolutes)\n\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%6.2f%%  %9.3f ms  n=%3d  avg %8.3f ms  %s\n" % (100 * v[1] / tot, v[1], v[0], v[1] / v[0], k))
    print(open(os.path.join(P, "%s_launch_summary.txt" % tag)).read())


# mangled-name fragments that pick the profiled template instance (the line table is per instance)
SYMS = {"narrowphase3_kernel": "narrowphase3_kernelILb0ELb0ELb1", "broadphase_grid_kernel": "broadphase_grid_kernelILb0ELb1"}


def kernel(rep, name, cubin):
    path = os.path.join(G, rep)
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    got = {}
    with open(os.path.join(P, "%s_%s_metrics.txt" % (tag, name)), "w") as f:
        f.write("ncu --set full --clock-control none --import-source on -k regex:%s -s 3 -c 1 ; command: python bench.py --steps 2 --warmup 3 --no-other-configs\n" % name)
        f.write("C5 workload, 2^20 pose-sets per launch (profiled launch: serialised, NOT a bench number)\n\n")
        for i, h in enumerate(hdr):
            if h in KEEP:
                f.write("%-75s %-16s %s\n" % (h, units[i], vals[i]))
                got[h] = (units[i], vals[i])
    env = dict(os.environ, NCU_CUBIN=cubin)
    hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot_lines.py"), path, SYMS.get(name, name + "ILb0"), "40"],
                         capture_output=True, text=True, env=env).stdout
    open(os.path.join(P, "%s_%s_hot_lines.txt" % (tag, name)), "w").write(hot)

    def tob(u, v):
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    return tob(*got["dram__bytes_read.sum"]) + tob(*got["dram__bytes_write.sum"])


launches()
# round 1 / first half of round 2: narrowphase_kernel + broadphase_kernel; from "r2f" on the defaults are
# narrowphase3_kernel (np3_kernels.cu) and broadphase_grid_kernel (split_kernels.cu)
NP, NPCUBIN, BP = (("narrowphase3_kernel", "np3_kernels", "broadphase_grid_kernel") if tag >= "r2f"
                   else ("narrowphase_kernel", "split_kernels", "broadphase_kernel"))
t = kernel("prof_%s_narrowphase.ncu-rep" % tag, NP, NPCUBIN)
kernel("prof_%s_broadphase.ncu-rep" % tag, BP, "split_kernels")
json.dump({"dram_bytes_per_launch": t, "kernel": NP,
           "source": "profiles/%s_%s_metrics.txt (ncu --set full, one launch on the C5 batch)" % (tag, NP)},
          open(os.path.join(P, "traffic.json"), "w"))
print("traffic", t)
