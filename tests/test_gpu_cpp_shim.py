"""The C++ drop-in headers (include/mpt_b200) driven like the reference's main.cpp: Map3D::safeEdge one edge
at a time == Map3D::safeEdges batched == the oracle's MeshHandler::isInCollision on the same edges."""
import os
import struct
import subprocess

import numpy as np
import pytest

from helpers import Scene, unit_box
from motionplanningtoolkit_b200 import meshio, scenes
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_map3d_safe_edges(built, tmp_path):
    exe = str(tmp_path / "omnidirectional_demo")
    libdir = os.path.join(ROOT, "motionplanningtoolkit_b200")
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "omnidirectional_demo.cpp"), "-o", exe,
                           "-L", libdir, "-lmptcc", "-Wl,-rpath," + libdir])
    box = os.path.join(meshio.DATA_DIR, "unit_box.mptm")
    out = str(tmp_path / "edges.bin")
    r = subprocess.run([exe, box, box, "400", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(out, "rb").read()
    n, = struct.unpack_from("<Q", raw, 0)
    assert n == 400
    b = unit_box()
    sc = Scene([b], [b])
    off, n_unsafe = 8, 0
    for _ in range(n):
        v = struct.unpack_from("<6d2B", raw, off)
        off += 50
        pts = scenes.omnidirectional_edge_poses(v[0:3], v[3:6], 0.1)
        want_safe = not O.is_in_collision(sc.oenv, sc.olinks, scenes.translation_poses(pts))
        assert bool(v[6]) == want_safe and bool(v[7]) == want_safe
        n_unsafe += not want_safe
    assert 0 < n_unsafe < n


def test_prm_roadmap_batched_equals_sequential(built, tmp_path):
    """BASELINE config 2 as a call-site change: the PRM roadmap built with one device call per
    slice of milestones equals the roadmap built with one Map3D::safeEdge per edge."""
    exe = str(tmp_path / "prm_roadmap")
    libdir = os.path.join(ROOT, "motionplanningtoolkit_b200")
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "prm_roadmap.cpp"), "-o", exe,
                           "-L", libdir, "-lmptcc", "-Wl,-rpath," + libdir])
    box = os.path.join(meshio.DATA_DIR, "unit_box.mptm")
    r = subprocess.run([exe, box, box, "300", "100"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "identical roadmaps: yes" in r.stdout
    print(r.stdout)


def test_is_in_collision_with_self_collision_flag(built, tmp_path):
    """SURVEY 8f N4 through the reference's own signature: isInCollision(env, agent, poses, checkSelfCollision)."""
    exe = str(tmp_path / "self_collision_check")
    libdir = os.path.join(ROOT, "motionplanningtoolkit_b200")
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "self_collision_check.cpp"), "-o", exe,
                           "-L", libdir, "-lmptcc", "-Wl,-rpath," + libdir])
    r = subprocess.run([exe, os.path.join(meshio.DATA_DIR, "unit_box.mptm")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.split() == ["apart", "0", "0", "touching", "0", "1", "overlapping", "0", "1", "just_apart", "0", "0"]


def _run_grid_demo(exe, tmp_path, env_name, location, merge):
    inst = str(tmp_path / "demo.inst")
    with open(inst, "w") as f:      # the reference's .inst syntax, read from a FILE by InstanceFileMap
        f.write("Agent Mesh ? %s\n" % os.path.join(meshio.DATA_DIR, "blimp.mptm" if merge else "unit_box.mptm"))
        f.write("Agent Merge Sub Meshes ? %s\n" % ("true" if merge else "false"))
        f.write("Environment Mesh ? %s\n" % os.path.join(meshio.DATA_DIR, env_name + ".mptm"))
        f.write("Environment Location ? %s\n" % " ".join(repr(float(x)) for x in location))
        f.write("Environment Bounding Box ? %s\n" % ("-40 40 -40 40 -30 30" if merge else "-3 3 -3 3 -3 3"))
        f.write("Grid Size ? %s\n" % ("10" if merge else "0.75"))
        f.write("Collision Check Delta t ? %s\n" % ("2.5" if merge else "0.2"))
    out = str(tmp_path / "demo.bin")
    r = subprocess.run([exe, inst, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout, open(out, "rb").read()


def _check_demo_sections(raw, env_meshes, agent_mesh, location):
    off, seen = 0, {}
    while off < len(raw):
        tag, fmt, links, n = struct.unpack_from("<IIIQ", raw, off)
        off += 20
        rec = 7 if fmt == 1 else 12
        poses = np.frombuffer(raw, dtype="<f8", count=n * links * rec, offset=off).reshape(n, links, rec)
        off += 8 * n * links * rec
        safe = np.frombuffer(raw, dtype=np.uint8, count=n, offset=off).astype(bool)
        off += n
        agent = agent_mesh if tag != 4 else meshio.agent_submesh(meshio.load_mptm("blimp"), "last") if agent_mesh is not None and len(agent_mesh[1]) > 100 else agent_mesh
        sc = Scene(env_meshes, [agent] * links, location)
        want = ~O.unpack_bits(sc.oracle_bits(poses, fmt), n)
        bad = np.nonzero(want != safe)[0]
        assert bad.size == 0, "section %d: %d of %d workspace verdicts differ from the oracle, first %s" % (tag, bad.size, n, bad[:5])
        seen[tag] = (n, int(safe.sum()))
    return seen


@pytest.mark.parametrize("flavour", ["plain", "reference", "patched"])
def test_grid_and_prmlite_call_sites(built, tmp_path, flavour):
    """safePoses (nested and flat), safePose, safePosesBatch, multi-link isInCollision, a file-based InstanceFileMap,
    a ROTATED "Environment Location" and mergeSubMeshes = true, all through the C++ shim; every pose-set the demo
    hands to the workspace is re-decided by the oracle.  `reference` / `patched` run the binaries the CPU suite built
    with the reference's GridDiscretization header (unchanged / with the batching patch applied)."""
    libdir = os.path.join(ROOT, "motionplanningtoolkit_b200")
    if flavour == "plain":
        exe = str(tmp_path / "grid_prmlite_demo")
        subprocess.check_call(["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"),
                               os.path.join(ROOT, "examples", "grid_prmlite_demo.cpp"), "-o", exe,
                               "-L", libdir, "-lmptcc", "-Wl,-rpath," + libdir])
    else:
        exe = os.path.join(ROOT, "examples", "bin", "grid_prmlite_demo_" + flavour)
        if not os.path.exists(exe):
            pytest.skip("examples/bin/grid_prmlite_demo_%s was not built (the CPU suite builds it where the reference exists)" % flavour)
    # unit boxes, environment rotated about a skew axis and shifted
    q = np.array([0.9, 0.1, -0.3, 0.3]); q /= np.linalg.norm(q)
    loc = np.concatenate([[0.4, -0.2, 0.3], q])
    text, raw = _run_grid_demo(exe, tmp_path, "unit_box", loc, merge=False)
    assert "all consistent: yes" in text, text
    if flavour != "plain":
        assert "consistent with the per-cell verdicts: yes" in text, text
    b = unit_box()
    seen = _check_demo_sections(raw, [b], b, loc)
    assert set(seen) == {1, 2, 3, 4} and all(0 < s < n for n, s in seen.values()), seen
    # the blimp with ALL its sub-meshes (mergeSubMeshes = true) in the model.dae environment, rotated as well
    loc2 = np.concatenate([[-60.0, -50.0, -40.0], q])
    text, raw = _run_grid_demo(exe, tmp_path, "model", loc2, merge=True)
    assert "all consistent: yes" in text, text
    env = meshio.environment_submeshes(meshio.load_mptm("model"))
    seen = _check_demo_sections(raw, env, meshio.agent_submesh(meshio.load_mptm("blimp"), "merge"), loc2)
    assert seen[1][1] < seen[1][0], seen


def test_speculative_rrt_reproduces_the_sequential_tree(built, tmp_path):
    """SURVEY 8f N1 for RRT (planners/rrt.hpp:42-94): K speculated iterations per device call grow exactly the tree
    of the one-safeEdge-per-iteration loop, with ~K times fewer collision calls."""
    exe = str(tmp_path / "rrt_speculative")
    libdir = os.path.join(ROOT, "motionplanningtoolkit_b200")
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "rrt_speculative.cpp"), "-o", exe,
                           "-L", libdir, "-lmptcc", "-Wl,-rpath," + libdir])
    box = os.path.join(meshio.DATA_DIR, "unit_box.mptm")
    for k in ("1", "16", "64"):
        r = subprocess.run([exe, box, box, "1500", k], capture_output=True, text=True)
        assert r.returncode == 0 and "identical trees: yes" in r.stdout, r.stdout + r.stderr
        print(r.stdout)
