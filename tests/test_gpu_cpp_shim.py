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
