#!/usr/bin/env python3
"""Pin tests/golden/*.npz against REAL FCL through oracle/_ref/fcl_harness (oracle/fcl_pin/README.md).

  make -C oracle _ref && python tools/pin_golden_with_fcl.py [--reference /root/reference] [--dry-run]

For every golden file whose scene the reference can load by itself (its own mesh files through its own
AssimpMeshLoader), the harness evaluates MeshHandler::isInCollision per pose-set; if the verdict words equal the
file's `bits`, the file is rewritten with pinned_by = "FCL via oracle/_ref/fcl_harness (<pkg-config version>)".
Differences are printed (pose indices) and the file is left untouched.  Also checks the sub-mesh split and order
AssimpMeshLoader::get produces against the .mptm fixtures (mesh-ingestion parity).
"""
import argparse
import glob
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from motionplanningtoolkit_b200 import meshio  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "fcl_harness")
MESH = {"unit_box": "mesh_models/agent_models/unit_box.dae", "blimp": "mesh_models/agent_models/blimp.3ds",
        "divider": "mesh_models/environment_models/divider.dae", "model": "mesh_models/environment_models/model.dae"}


def fnv1a(mesh):
    h = 1469598103934665603
    data = np.ascontiguousarray(mesh[0], dtype="<f8").tobytes() + np.ascontiguousarray(mesh[1], dtype="<u4").tobytes()
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def write_obj(path, mesh):
    """One un-grouped OBJ object: Assimp (flags 0) turns it into ONE aiMesh, i.e. one sub-mesh for the handlers."""
    with open(path, "w") as f:
        for v in mesh[0]:
            f.write("v %.17g %.17g %.17g\n" % tuple(v))
        for t in mesh[1]:
            f.write("f %d %d %d\n" % (t[0] + 1, t[1] + 1, t[2] + 1))


def scene_of(name, ref, tmp):
    """-> (env file, agent file, n_links, self) for a golden file name"""
    env = "unit_box" if name.startswith(("c1_", "c4_", "n4_")) else ("divider" if "divider" in name else "model")
    links = 11 if name.startswith("c4_") else 4 if name.startswith("n4_") else 1
    if name.startswith(("c1_", "c4_", "n4_")):
        agent = os.path.join(ref, MESH["unit_box"])
    elif "_last_" in name:
        agent = os.path.join(ref, MESH["blimp"])          # the reference keeps the last sub-mesh by itself
    else:
        agent = os.path.join(tmp, "blimp_merged.obj")      # "merge" is not something the reference can do: one OBJ object
        if not os.path.exists(agent):
            write_obj(agent, meshio.agent_submesh(meshio.load_mptm("blimp"), "merge"))
    return os.path.join(ref, MESH[env]), agent, links


def check_meshes(ref):
    ok = True
    for name, rel in MESH.items():
        out = subprocess.run([HARNESS, "--dump-meshes", os.path.join(ref, rel)], capture_output=True, text=True).stdout.split("\n")
        subs = meshio.load_mptm(name)
        got = [tuple(l.split()) for l in out[1:1 + int(out[0])]]
        want = [(str(len(v)), str(len(t)), fnv1a((v, t))) for v, t in subs]
        same = got == want
        ok &= same
        print("mesh %-9s %3d sub-meshes from Assimp, %3d in the fixture: %s" % (name, len(got), len(want), "identical" if same else "DIFFERENT"))
        if not same:
            for i, (g, w) in enumerate(zip(got, want)):
                if g != w:
                    print("   sub-mesh %d: assimp nv/nt/digest %s vs fixture %s" % (i, g, w))
                    break
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--dry-run", action="store_true")
    args = ap.parse_args()
    if not os.path.exists(HARNESS):
        raise SystemExit("oracle/_ref/fcl_harness is not built: needs FCL 0.3.2, Assimp and Boost (make -C oracle _ref)")
    ver = subprocess.run(["pkg-config", "--modversion", "fcl"], capture_output=True, text=True).stdout.strip() or "?"
    check_meshes(args.reference)
    tmp = tempfile.mkdtemp()
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
        name = os.path.basename(path)
        g = dict(np.load(path))
        env, agent, links = scene_of(name, args.reference, tmp)
        poses = np.ascontiguousarray(g["poses"], dtype="<f8")
        n = poses.shape[0]
        pin, pout = os.path.join(tmp, "poses.bin"), os.path.join(tmp, "out.bits")
        with open(pin, "wb") as f:
            f.write(b"MPTPIN01" + struct.pack("<IIQ", int(g["fmt"]), links, n) + poses.tobytes())
        pose7 = " ".join(repr(float(x)) for x in np.asarray(g["pose7"]).ravel()[:7])
        cmd = [HARNESS, env, pose7, agent, pin, pout] + (["--self"] if int(g.get("self_collision", 0)) else [])
        rc = subprocess.run(cmd).returncode
        if rc != 0:
            print("%-40s harness failed (rc %d)" % (name, rc))
            continue
        got = np.fromfile(pout, dtype="<u4")
        want = np.asarray(g["bits"], dtype=np.uint32)
        if np.array_equal(got, want):
            print("%-40s %6d pose-sets: FCL == oracle -> pinned" % (name, n))
            if not args.dry_run:
                g["pinned_by"] = np.array("FCL %s via oracle/_ref/fcl_harness (reference meshhandler.hpp unmodified)" % ver)
                np.savez_compressed(path, **g)
        else:
            unp = lambda w: np.unpackbits(w.view(np.uint8), bitorder="little")[:n]
            bad = np.nonzero(unp(got) != unp(want))[0]
            print("%-40s %6d pose-sets: %d DIFFER from FCL, first %s -- file left unpinned" % (name, n, bad.size, bad[:10]))


if __name__ == "__main__":
    main()
