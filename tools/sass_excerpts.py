#!/usr/bin/env python3
"""SASS excerpts of the default narrowphase kernel (np3, quantised nodes) for profiles/: opcode counts and the
instruction groups that show what the kernel is built from.  usage: sass_excerpts.py <round-tag>"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2g"
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "motionplanningtoolkit_b200", "libmptcc.so")], cwd=tmp,
                      stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.startswith("np3_kernels")][0]
sass = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
out = ["SASS of the third-generation narrowphase (nvdisasm of the np3_kernels cubin inside motionplanningtoolkit_b200/libmptcc.so, sm_100a)", ""]
for inst, what in (("ILb0ELb0ELb1", "default: 32-byte quantised nodes"), ("ILb0ELb1ELb0", "agent_top = n: fp32 nodes, agent tree staged by TMA")):
    m = re.search(r"\.text\.(\S*narrowphase3_kernel%s\S*):" % inst, sass)
    if not m:
        continue
    i = m.end()
    j = sass.find("//---------------------", i)
    lines = [l for l in sass[i:j if j > 0 else None].splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    ops = collections.Counter()
    for l in lines:
        mm = re.search(r"\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", l)
        if mm:
            ops[mm.group(1).split(".")[0]] += 1
    out.append("== %s  (%s)" % (m.group(1), what))
    out.append("   %d instructions: " % len(lines) + ", ".join("%s %d" % (k, ops[k]) for k in
               ("FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "FSETP", "LOP3", "PRMT", "SHF", "LDG", "LDGSTS", "LDS", "STS", "LDL", "STL", "UBLKCP", "SYNCS")))

    def grp(title, pat, limit):
        sel = [l for l in lines if re.search(pat, l)][:limit]
        if sel:
            out.append("   -- " + title)
            out.extend("   " + l.rstrip() for l in sel)
    grp("cp.async prefetch of the next candidates' records", r"LDGSTS", 6)
    grp("256-bit node gathers", r"LDG\.E\.(?:\w+\.)*256", 6)
    grp("quantised-node decode: mask, subnormal operand times 2^127 minus 0.5; link by PRMT", r"0x3fffff|PRMT.*0xf732|1\.7014118346046923173e\+38", 14)
    grp("two-wide SAT (first 12)", r"FFMA2|FMUL2|FADD2", 12)
    grp("TMA bulk copy of the agent tree and its mbarrier", r"UBLKCP|SYNCS", 8)
    out.append("")
open(os.path.join(ROOT, "profiles", "%s_np3_sass_excerpts.txt" % tag), "w").write("\n".join(out))
print("\n".join(out)[:3000])
