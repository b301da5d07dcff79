#!/usr/bin/env python3
"""Convert the reference's bundled mesh fixtures to the flat ``.mptm`` format.

Run once in the build container (needs /root/reference); the outputs under
``motionplanningtoolkit_b200/data/`` are committed so nothing on the GPU box reads
/root/reference.

The conversion reproduces what ``AssimpMeshLoader::get``
(/root/reference/utilities/assimp_mesh_loader.hpp:20-60) hands to FCL when Assimp
is called with flags 0 (``ReadFile(pFile, 0)``, assimp_mesh_loader.hpp:12):

* one sub-mesh per COLLADA ``<triangles>`` group, in scene-instance order, and one
  sub-mesh per 3DS material split, in material order;
* vertices are un-indexed (three fresh vertices per face, no JoinIdenticalVertices);
* raw mesh-local float32 coordinates (node matrices, ``<unit>``, ``up_axis`` ignored);
* faces that are not triangles (``<lines>``) are dropped (assimp_mesh_loader.hpp:50-53)
  and sub-meshes left without triangles are skipped by the mesh handlers
  (/root/reference/utilities/meshhandler.hpp:42,128), so they are not emitted.

File format (little endian): ``b"MPTM"``, u32 version=1, u32 n_submeshes, then per
sub-mesh: u32 nv, u32 nt, f64 xyz[nv][3], u32 tri[nt][3].
"""
import os
import re
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np

REF = "/root/reference/mesh_models"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                   "motionplanningtoolkit_b200", "data")


def write_mptm(path, submeshes):
    with open(path, "wb") as f:
        f.write(b"MPTM")
        f.write(struct.pack("<II", 1, len(submeshes)))
        for name, xyz, tri in submeshes:
            xyz = np.ascontiguousarray(xyz, dtype=np.float64)
            tri = np.ascontiguousarray(tri, dtype=np.uint32)
            f.write(struct.pack("<II", xyz.shape[0], tri.shape[0]))
            f.write(xyz.tobytes())
            f.write(tri.tobytes())


def unindex(pos, faces):
    """Three fresh vertices per face, faces (0,1,2),(3,4,5),..."""
    faces = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
    xyz = pos[faces.reshape(-1)]
    tri = np.arange(faces.shape[0] * 3, dtype=np.uint32).reshape(-1, 3)
    return xyz, tri


def load_dae(path):
    ns = {"c": "http://www.collada.org/2005/11/COLLADASchema"}
    root = ET.parse(path).getroot()
    geoms = {}
    for g in root.findall(".//c:library_geometries/c:geometry", ns):
        geoms[g.get("id")] = g
    libnodes = {}
    for n in root.iter("{%s}node" % ns["c"]):
        if n.get("id"):
            libnodes[n.get("id")] = n
    order = []

    def visit(node):
        # Assimp ColladaLoader::BuildHierarchy order: real child nodes, then resolved
        # <instance_node>s, then this node's own <instance_geometry> meshes.
        for ch in node.findall("c:node", ns):
            visit(ch)
        for inst in node.findall("c:instance_node", ns):
            visit(libnodes[inst.get("url").lstrip("#")])
        for inst in node.findall("c:instance_geometry", ns):
            gid = inst.get("url").lstrip("#")
            if gid not in order:      # Assimp caches one aiMesh per (geometry, sub-mesh, material)
                order.append(gid)

    for vs in root.findall(".//c:library_visual_scenes/c:visual_scene", ns):
        for n in vs.findall("c:node", ns):
            visit(n)
    subs = []
    for gid in order:
        mesh = geoms[gid].find("c:mesh", ns)
        sources = {}
        for s in mesh.findall("c:source", ns):
            fa = s.find("c:float_array", ns)
            vals = np.array([np.float32(float(t)) for t in fa.text.split()], dtype=np.float32)
            stride = int(s.find("c:technique_common/c:accessor", ns).get("stride", "3"))
            sources[s.get("id")] = vals.reshape(-1, stride)
        verts = {}
        for v in mesh.findall("c:vertices", ns):
            for inp in v.findall("c:input", ns):
                if inp.get("semantic") == "POSITION":
                    verts[v.get("id")] = sources[inp.get("source").lstrip("#")]
        for prim in mesh:
            tag = prim.tag.split("}")[1]
            if tag != "triangles":
                continue            # <lines>: 2-index faces, all dropped -> empty sub-mesh -> skipped
            inputs = prim.findall("c:input", ns)
            nin = 1 + max(int(i.get("offset")) for i in inputs)
            voff = None
            pos = None
            for i in inputs:
                if i.get("semantic") == "VERTEX":
                    voff = int(i.get("offset"))
                    pos = verts[i.get("source").lstrip("#")]
            p = np.array(prim.find("c:p", ns).text.split(), dtype=np.int64).reshape(-1, nin)
            idx = p[:, voff]
            assert idx.size == 3 * int(prim.get("count"))
            xyz, tri = unindex(pos[:, :3].astype(np.float64), idx)
            subs.append((gid, xyz, tri))
    return subs


def load_3ds(path):
    data = open(path, "rb").read()
    materials = []
    objects = []

    def cstr(off):
        end = data.index(b"\0", off)
        return data[off:end].decode("latin1"), end + 1

    def walk(off, end, ctx):
        while off + 6 <= end:
            cid, clen = struct.unpack_from("<HI", data, off)
            body, cend = off + 6, off + clen
            if cid in (0x4D4D, 0x3D3D):
                walk(body, cend, ctx)
            elif cid == 0xAFFF:
                m = {}
                walk(body, cend, m)
                materials.append(m.get("name", ""))
            elif cid == 0xA000:
                ctx["name"], _ = cstr(body)
            elif cid == 0x4000:
                name, nb = cstr(body)
                o = {"name": name, "verts": np.zeros((0, 3), np.float32), "faces": [], "facemat": []}
                walk(nb, cend, o)
                objects.append(o)
            elif cid == 0x4100:
                walk(body, cend, ctx)
            elif cid == 0x4110:
                n, = struct.unpack_from("<H", data, body)
                ctx["verts"] = np.frombuffer(data, dtype="<f4", count=3 * n, offset=body + 2).reshape(n, 3)
            elif cid == 0x4120:
                n, = struct.unpack_from("<H", data, body)
                f = np.frombuffer(data, dtype="<u2", count=4 * n, offset=body + 2).reshape(n, 4)
                ctx["faces"] = f[:, :3].astype(np.int64)
                walk(body + 2 + 8 * n, cend, ctx)
            elif cid == 0x4130:
                mname, nb = cstr(body)
                n, = struct.unpack_from("<H", data, nb)
                ids = np.frombuffer(data, dtype="<u2", count=n, offset=nb + 2).astype(np.int64)
                ctx["facemat"].append((mname, ids))
            off = cend

    walk(0, len(data), {})
    subs = []
    for o in objects:
        nf = len(o["faces"])
        if nf == 0 or len(o["verts"]) == 0:
            continue                                  # Assimp drops empty objects
        fm = np.full(nf, -1, dtype=np.int64)
        for mname, ids in o["facemat"]:
            fm[ids] = materials.index(mname)
        # Assimp splits one mesh per used material, in material order; faces without a
        # material get the default material, appended last.
        for mi in list(range(len(materials))) + [-1]:
            sel = np.nonzero(fm == mi)[0]
            if sel.size == 0:
                continue
            xyz, tri = unindex(o["verts"].astype(np.float64), o["faces"][sel])
            subs.append(("%s/%s" % (o["name"], materials[mi] if mi >= 0 else "default"), xyz, tri))
    return subs


def main():
    os.makedirs(OUT, exist_ok=True)
    jobs = [
        ("agent_models/unit_box.dae", "unit_box.mptm", load_dae),
        ("agent_models/blimp.3ds", "blimp.mptm", load_3ds),
        ("environment_models/model.dae", "model.mptm", load_dae),
        ("environment_models/divider.dae", "divider.mptm", load_dae),
    ]
    for src, dst, fn in jobs:
        subs = fn(os.path.join(REF, src))
        write_mptm(os.path.join(OUT, dst), subs)
        nt = sum(s[2].shape[0] for s in subs)
        allv = np.concatenate([s[1] for s in subs])
        print("%-34s -> %-14s submeshes=%4d tris=%6d bbox=[%s .. %s]" % (
            src, dst, len(subs), nt, allv.min(0).round(3), allv.max(0).round(3)))
        if len(subs) <= 8:
            for s in subs:
                print("      %-30s nv=%5d nt=%5d" % (s[0], s[1].shape[0], s[2].shape[0]))


if __name__ == "__main__":
    sys.exit(main())
