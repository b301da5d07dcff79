"""Flat mesh fixtures (``.mptm``) and the sub-mesh selection rules of the reference's mesh handlers.

``load_mptm`` returns the per-sub-mesh (xyz float64 [nv,3], tri uint32 [nt,3]) lists that
``AssimpMeshLoader::get`` (/root/reference/utilities/assimp_mesh_loader.hpp:20-60) would hand over;
the converter is ``tools/convert_meshes.py``.
"""
import os
import struct

import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def load_mptm(name_or_path):
    path = name_or_path
    if not os.path.isabs(path) and not os.path.exists(path):
        path = os.path.join(DATA_DIR, name_or_path if name_or_path.endswith(".mptm") else name_or_path + ".mptm")
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:4] != b"MPTM":
        raise ValueError("%s: not an MPTM file" % path)
    ver, nsub = struct.unpack_from("<II", buf, 4)
    if ver != 1:
        raise ValueError("%s: unsupported MPTM version %d" % (path, ver))
    off = 12
    subs = []
    for _ in range(nsub):
        nv, nt = struct.unpack_from("<II", buf, off)
        off += 8
        xyz = np.frombuffer(buf, dtype="<f8", count=3 * nv, offset=off).reshape(nv, 3).copy()
        off += 24 * nv
        tri = np.frombuffer(buf, dtype="<u4", count=3 * nt, offset=off).reshape(nt, 3).copy()
        off += 12 * nt
        subs.append((xyz, tri))
    return subs


def environment_submeshes(subs):
    """StaticEnvironmentMeshHandler: one collision object per non-empty sub-mesh
    (/root/reference/utilities/meshhandler.hpp:38-50)."""
    return [(v, t) for (v, t) in subs if len(v) and len(t)]


def agent_submesh(subs, mode="last"):
    """SimpleAgentMeshHandler keeps only the LAST non-empty sub-mesh: ``agentModel`` is overwritten
    in the loop (/root/reference/utilities/meshhandler.hpp:124-134).  ``mode="merge"`` is the
    evidently intended behaviour (all sub-meshes in one model via repeated addSubModel)."""
    ne = environment_submeshes(subs)
    if not ne:
        raise ValueError("agent mesh has no non-empty sub-mesh")
    if mode == "last":
        return ne[-1]
    if mode == "merge":
        vs, ts, base = [], [], 0
        for v, t in ne:
            vs.append(v)
            ts.append(t + base)
            base += len(v)
        return np.concatenate(vs), np.concatenate(ts).astype(np.uint32)
    if isinstance(mode, int):
        return ne[mode]
    raise ValueError("agent submesh mode must be 'last', 'merge' or an index")
