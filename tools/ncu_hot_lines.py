#!/usr/bin/env python3
"""Join ncu's per-SASS-instruction stall samples with nvdisasm line info -> hottest source lines.

usage: ncu_hot_lines.py <report.ncu-rep> [kernel-substring] [top-N]
The cubin is extracted from motionplanningtoolkit_b200/libmptcc.so (must be the build that was profiled);
NCU_CUBIN=split_kernels|collide_kernel picks the translation unit, NCU_SRC the file whose lines are shown
(lines of other files of the unit are attributed to the same numbers, so read them with care).
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
rep = sys.argv[1]
ksub = sys.argv[2] if len(sys.argv) > 2 else "collide_kernelILb0"
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40

tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(os.path.join(ROOT, "motionplanningtoolkit_b200", "libmptcc.so"))],
                      cwd=tmp, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.startswith(os.environ.get("NCU_CUBIN", "collide_kernel"))][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
# offset -> (function, line)
line_of = {}
cur_fn, cur_line = None, ("?", 0)
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        cur_fn = m.group(1)
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur_fn:
        line_of[(cur_fn, int(m.group(1), 16))] = cur_line

_src_cache = {}


def src_text(f, ln):
    if f not in _src_cache:
        path = os.path.join(ROOT, "motionplanningtoolkit_b200", "csrc", f)
        _src_cache[f] = open(path).read().splitlines() if os.path.exists(path) else []
    t = _src_cache[f]
    return t[ln - 1].strip()[:80] if 0 < ln <= len(t) else "?"


SORT = os.environ.get("NCU_SORT", "samples")  # or "inst"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the report may hold several kernels; each starts with a "Kernel Name" row followed by a header row
blocks, i = [], 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr):
                body.append(rows[j])
            j += 1
        blocks.append((name, hdr, body))
        i = j
    else:
        i += 1
name, hdr, body = blocks[0]
ci = {h: k for k, h in enumerate(hdr)}
fn = [f for (f, _o) in line_of if ksub in f]
fn = fn[0] if fn else None
base = int(body[0][ci["Address"]], 16)
stall_keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in body:
    off = int(r[ci["Address"]], 16) - base
    ln = line_of.get((fn, off), ("?", -1))
    a = agg[ln]
    a["samples"] += int(r[ci["# Samples"]])
    a["inst"] += int(r[ci["Instructions Executed"]])
    a["thr"] += int(r[ci["Thread Instructions Executed"]])
    for k in stall_keys:
        a[k] += int(r[ci[k]] or 0)
    tot["samples"] += int(r[ci["# Samples"]])
    tot["inst"] += int(r[ci["Instructions Executed"]])
    tot["thr"] += int(r[ci["Thread Instructions Executed"]])
    for k in stall_keys:
        tot[k] += int(r[ci[k]] or 0)
print("kernel:", name)
print("total samples %d, warp-instructions %d, avg threads/inst %.1f" % (tot["samples"], tot["inst"], tot["thr"] / max(1, tot["inst"])))
print("stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * tot[k] / max(1, tot["samples"]))
                              for k in sorted(stall_keys, key=lambda k: -tot[k])[:8]))
print("%7s %7s %5s  line  source" % ("samp%", "inst%", "thr"))
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][SORT])[:topn]:
    st = sorted(((a[k], k[6:]) for k in stall_keys), reverse=True)[:3]
    text = src_text(*ln)
    print("%6.2f%% %6.2f%% %5.1f  %-22s %-80s %s" % (100.0 * a["samples"] / max(1, tot["samples"]), 100.0 * a["inst"] / max(1, tot["inst"]),
                                                  a["thr"] / max(1, a["inst"]), "%s:%d" % (ln[0][:16], ln[1]), text,
                                                  " ".join("%s:%d" % (k, v) for v, k in st if v)))
