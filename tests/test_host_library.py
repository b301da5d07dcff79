"""CPU tests of libmptcc's host side: the C ABI surface, the host BVH builder (must be bit-identical to
the oracle's restatement of FCL's builder), failure modes without a GPU, and the sharding helpers.

Device code without a device: the arithmetic of the kernels -- pose decoding, object boxes, relative transform, candidate
record, the exact fp64 OBB and triangle tests, the fp32 OBB and triangle filters, the quantised-node fast pass -- lives in
headers (csrc/collide_device.cuh between [host-test:begin] / [host-test:end], csrc/np_sat.cuh) that tests/host_filters.cpp
and tests/host_sat.cpp compile for the CPU with stand-ins for the CUDA types; those functions are compared with the oracle
here (bit for bit where they restate FCL, for one-sidedness where they are conservative filters), and the broadphase grid
and cluster tree are built by a context without devices and checked for completeness.  What only runs on the GPU -- the
traversal itself -- is covered by the -m gpu tests."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from motionplanningtoolkit_b200 import capi, meshio, scenes, sharding
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "mptcc.h")).read()
    declared = set(re.findall(r"\b(mptcc_[a-z0-9_]+)\s*\(", header))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    nm = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (mptcc_[a-z0-9_]+)", nm))
    assert declared <= exported


def test_no_cpu_fallback(built):
    """Without a usable sm_100 device the context cannot be created: there is no CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.MptccError) as e:
        capi.Checker(1)
    assert e.value.code == capi.ERR_NO_DEVICE


def test_host_bvh_is_bit_identical_to_oracle(built):
    for name in ("unit_box", "blimp", "model", "divider"):
        for k, (v, t) in enumerate(meshio.environment_submeshes(meshio.load_mptm(name))):
            a, b = capi.HostBvh(v, t), O.Model(v, t)
            na, nb = a.nodes(), b.nodes()
            for key in na:
                assert np.array_equal(na[key], nb[key]), (name, k, key)
            assert np.array_equal(a.primitive_indices(), b.primitive_indices())
            (ba, ca, ra), (bb, cb, rb) = a.local_aabb(), b.local_aabb()
            assert np.array_equal(ba, bb) and np.array_equal(ca, cb) and ra == rb


def test_host_bvh_on_procedural_meshes(built):
    env = scenes.synthetic_environment(12)
    for v, t in env:
        a, b = capi.HostBvh(v, t), O.Model(v, t)
        na, nb = a.nodes(), b.nodes()
        for key in na:
            assert np.array_equal(na[key], nb[key]), key
        assert a.num_nodes == 2 * len(t) - 1 and a.depth >= int(np.ceil(np.log2(len(t)))) + 1


def _half_to_float(h):
    return np.asarray(h, np.uint16).view(np.float16).astype(np.float64)


@pytest.mark.parametrize("breadth_first", [False, True])
def test_quantised_nodes_decode_within_their_bounds(built, breadth_first):
    """The 32-byte traversal nodes (csrc/device_layout.hpp, nodeq), decoded here the way np_sat.cuh decodes them, against
    the fp64 nodes they stand for: link round trip (24-bit two's complement, device numbering), axes within 2e-6 of
    FCL's (third axis possibly negated), centre = the fp32 rounding of the recentred centre, extents widened as DESIGN.md
    section 6 states and never by more than binary16 costs, and the decoded box CONTAINS FCL's box."""
    meshes = [(n, v, t) for n in ("blimp", "model", "divider") for (v, t) in meshio.environment_submeshes(meshio.load_mptm(n))]
    meshes += [("synthetic", v, t) for v, t in scenes.synthetic_environment(6)]
    for name, v, t in meshes:
        hb = capi.HostBvh(v, t)
        words, dev, ctr, ok = hb.quantised_nodes(breadth_first)
        assert ok, name
        nd = hb.nodes()
        n = hb.num_nodes
        w = words[dev]                                     # row i = the device node of FCL node i
        q0, q1 = w[:, :4].astype(np.uint64), w[:, 4:].astype(np.uint64)
        # link: bytes 2, 3 of q1.x and byte 3 of q1.w, sign-extended
        raw = (q1[:, 0] >> 16) | ((q1[:, 3] >> 24) << 16)
        link = np.where(raw & 0x800000, raw.astype(np.int64) - (1 << 24), raw.astype(np.int64))
        fc = nd["first_child"].astype(np.int64)
        want = np.where(fc >= 0, dev[np.maximum(fc, 0)].astype(np.int64), fc)
        assert np.array_equal(link, want), name
        if breadth_first:                                  # siblings adjacent, parents before children
            inner = fc >= 0
            assert np.all(dev[fc[inner] + 1] == dev[fc[inner]] + 1) and np.all(dev[fc[inner]] > dev[np.nonzero(inner)[0]])
        # quaternion: four 22-bit components in bits [0,22) [22,44) [44,66) [66,88) of q1.y, .z, .w
        bits = q1[:, 1] | (q1[:, 2] << 32)                 # low 64 of the 96
        Q = np.stack([bits & 0x3fffff, (bits >> 22) & 0x3fffff, ((bits >> 44) | (q1[:, 3] << 20)) & 0x3fffff,
                      (q1[:, 3] >> 2) & 0x3fffff], 1)
        # the kernel's evaluation: the bits as a subnormal float times 2^127 minus 0.5 = half the component, exactly
        half = (Q.astype(np.uint32).view(np.float32) * np.float32(2.0 ** 127) - np.float32(0.5)).astype(np.float64)
        qd = Q.astype(np.float64) / 2097152.0 - 1.0
        assert np.array_equal(half * 2.0, qd), name
        qw, qx, qy, qz = qd.T
        R = np.stack([np.stack([1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qw * qz), 2 * (qx * qz + qw * qy)], 1),
                      np.stack([2 * (qx * qy + qw * qz), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qw * qx)], 1),
                      np.stack([2 * (qx * qz - qw * qy), 2 * (qy * qz + qw * qx), 1 - 2 * (qx * qx + qy * qy)], 1)], 1)
        A = nd["axis"].copy()                              # rows = box axes
        A[np.linalg.det(A) < 0, 2] *= -1.0                 # a left-handed triple is stored with axis 2 negated: the same box
        assert np.abs(R - A).max() <= 2e-6, (name, np.abs(R - A).max())
        # centre
        c32 = w[:, :3].copy().view(np.float32).astype(np.float64)
        assert np.array_equal(c32, (nd["To"] - ctr).astype(np.float32).astype(np.float64)), name
        # extents: binary16, rounded up from (e + sqrt(3) 4e-6 sum(e)) (1 + 1e-6)
        E = nd["extent"]
        eq = np.stack([_half_to_float(q0[:, 3] & 0xffff), _half_to_float(q0[:, 3] >> 16), _half_to_float(q1[:, 0] & 0xffff)], 1)
        lo = (E + np.sqrt(3.0) * 4e-6 * E.sum(1, keepdims=True)) * (1 + 1e-6)
        assert np.all(eq >= lo), name
        assert np.all(eq <= lo * (1 + 2.0 ** -10) * (1 + 1e-6) + 6.2e-5), name       # (one binary16 ulp; subnormal halves below 6.1e-5)
        # containment: every corner of FCL's box lies inside the decoded box
        for sx in (-1, 1):
            for sy in (-1, 1):
                for sz in (-1, 1):
                    corner = (sx * E[:, 0, None] * nd["axis"][:, 0] + sy * E[:, 1, None] * nd["axis"][:, 1] +
                              sz * E[:, 2, None] * nd["axis"][:, 2])     # relative to the (common) centre
                    proj = np.abs(np.einsum("nij,nj->ni", R, corner))
                    assert np.all(proj <= eq), name


def test_quantised_nodes_refuse_what_they_cannot_represent(built):
    """Extents beyond binary16 (65 000): the pack reports that it cannot be quantised (the kernels then traverse fp32 nodes)."""
    v, t = meshio.load_mptm("unit_box")[0]
    assert capi.HostBvh(v, t).quantised_nodes()[3]
    assert not capi.HostBvh(np.asarray(v) * 2.0e5, t).quantised_nodes()[3]


def _host_sat(tmp_path):
    """tests/host_sat.cpp: the arithmetic of csrc/np_sat.cuh compiled for the CPU (stand-ins for the CUDA types only)."""
    here = os.path.dirname(os.path.abspath(__file__))
    so = str(tmp_path / "host_sat.so")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-shared", "-fPIC",
                           "-I", os.path.join(here, "..", "motionplanningtoolkit_b200", "csrc"), os.path.join(here, "host_sat.cpp"), "-o", so])
    L = ctypes.CDLL(so)
    u32p, f32p = ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_float)
    L.host_pair_disjoint_q.restype = ctypes.c_uint
    L.host_pair_disjoint_q.argtypes = [u32p, u32p, u32p, f32p, ctypes.c_int]
    L.host_decode_difference.restype = ctypes.c_float
    L.host_decode_difference.argtypes = [u32p, u32p]
    L.host_pair_disjoint_f32.restype = ctypes.c_uint
    L.host_pair_disjoint_f32.argtypes = [f32p, f32p, f32p, f32p, ctypes.c_int]
    L.host_child_disjoint.restype = ctypes.c_uint
    L.host_child_disjoint.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, f32p, ctypes.c_int, ctypes.c_int]
    return L


def test_fp32_obb_filter_is_one_sided_against_the_oracle(built, tmp_path):
    """The narrowphase's fast pass (quantised nodes, child-frame two-wide SAT on 11 axes, margins of DESIGN.md section 6),
    compiled for the CPU from the kernel's own header, against FCL's fp64 obbDisjoint as the oracle restates it: on box
    pairs pushed to the edge of touching, "certainly disjoint" must never be said of a pair FCL finds overlapping -- in
    either descent direction -- and the filter must still catch nearly all pairs that are clearly apart.  The same for the
    variants the options select: the two-wide test on 64-byte fp32 nodes (qnode = 0) and the scalar 15-axis test of the np2
    kernel on either node format."""
    L = _host_sat(tmp_path)
    u32 = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
    f32 = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    rng = np.random.default_rng(20260)
    av, at = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    envs = sorted(meshio.environment_submeshes(meshio.load_mptm("divider")), key=lambda m: -len(m[1]))[:2] + scenes.synthetic_environment(2)
    hbA = capi.HostBvh(av, at)
    wA, devA, ctrA, okA = hbA.quantised_nodes(True)
    ndA = hbA.nodes()
    n32A = _device_model(av, at)["node32"]
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    wrong = caught = apart = checked = wrong_variants = 0
    for ev, et in envs:
        hbE = capi.HostBvh(ev, et)
        n32E = _device_model(ev, et)["node32"]
        wE, devE, ctrE, okE = hbE.quantised_nodes(False)
        ndE = hbE.nodes()
        assert okA and okE
        # the scalar and the two-wide decode give the same boxes
        for _ in range(200):
            a, b = rng.integers(hbE.num_nodes), rng.integers(hbA.num_nodes)
            # (to within the rounding of a*b+c as FMA or not: the two-wide decode spells its FMAs out, the scalar one leaves
            # the contraction to the compiler)
            assert L.host_decode_difference(u32(wE[devE[a]].copy()), u32(wA[devA[b]].copy())) <= 2.4e-7
        for _ in range(400):
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            R = scenes.quat_to_matrix(q[None])[0]
            i, i2 = rng.integers(hbE.num_nodes, size=2)
            j, j2 = rng.integers(hbA.num_nodes, size=2)
            d = rng.normal(size=3); d /= np.linalg.norm(d)
            bE = lambda k: (ndE["axis"][k], ndE["To"][k], ndE["extent"][k])
            bA = lambda k: (ndA["axis"][k], ndA["To"][k], ndA["extent"][k])
            base = ndE["To"][i] - R @ ndA["To"][j]            # centres of (i, j) coincide at t = 0
            hi = 2.0 * (np.linalg.norm(ndE["extent"][i]) + np.linalg.norm(ndA["extent"][j])) + 1e-3
            lo = 0.0
            if not O.obb_overlap(R, base, bE(i), bA(j)) or O.obb_overlap(R, base + hi * d, bE(i), bA(j)):
                continue
            for _ in range(60):                               # the edge of touching along d, to the last bit
                mid = 0.5 * (lo + hi)
                if O.obb_overlap(R, base + mid * d, bE(i), bA(j)):
                    lo = mid
                else:
                    hi = mid
            for t in (lo, lo * (1 - 1e-9), lo * (1 - 1e-6), lo * (1 - 1e-3), hi, hi * (1 + 1e-6), hi * (1 + 1e-3), hi * 1.05 + 1e-3, hi * 1.5 + 1e-2):
                T = base + t * d
                Tc = R @ ctrA + T - ctrE
                rec = np.concatenate([R.ravel(), Tc, -(R.T @ Tc), [0.0]]).astype(np.float32)
                # children = environment nodes i, i2 against the agent node j; then children = agent nodes j, j2 against i
                m_env = L.host_pair_disjoint_q(u32(wA[devA[j]].copy()), u32(wE[devE[i]].copy()), u32(wE[devE[i2]].copy()), f32(rec), 1)
                m_ag = L.host_pair_disjoint_q(u32(wE[devE[i]].copy()), u32(wA[devA[j]].copy()), u32(wA[devA[j2]].copy()), f32(rec), 0)
                truth = {}
                for said_disjoint, (x, y) in ((m_env & 1, (i, j)), (m_env & 2, (i2, j)), (m_ag & 1, (i, j)), (m_ag & 2, (i, j2))):
                    overlap = truth[(x, y)] = O.obb_overlap(R, T, bE(x), bA(y))
                    checked += 1
                    wrong += bool(said_disjoint) and overlap
                    if (x, y) == (i, j) and t >= hi * 1.05:
                        apart += 1
                        caught += bool(said_disjoint)
                # fp32 nodes two-wide, then the scalar 15-axis test on fp32 and on quantised nodes
                sA, sE = np.ascontiguousarray(n32A[j]), np.ascontiguousarray(n32E[i])
                cE, cE2 = np.ascontiguousarray(n32E[i]), np.ascontiguousarray(n32E[i2])
                cA, cA2 = np.ascontiguousarray(n32A[j]), np.ascontiguousarray(n32A[j2])
                qa, qa2, qe, qe2 = (wA[devA[j]].copy(), wA[devA[j2]].copy(), wE[devE[i]].copy(), wE[devE[i2]].copy())
                for m1, m2 in ((L.host_pair_disjoint_f32(f32(sA), f32(cE), f32(cE2), f32(rec), 1), L.host_pair_disjoint_f32(f32(sE), f32(cA), f32(cA2), f32(rec), 0)),
                               (L.host_child_disjoint(vp(sA), vp(cE), vp(cE2), f32(rec), 1, 0), L.host_child_disjoint(vp(sE), vp(cA), vp(cA2), f32(rec), 0, 0)),
                               (L.host_child_disjoint(vp(qa), vp(qe), vp(qe2), f32(rec), 1, 1), L.host_child_disjoint(vp(qe), vp(qa), vp(qa2), f32(rec), 0, 1))):
                    for said_disjoint, key in ((m1 & 1, (i, j)), (m1 & 2, (i2, j)), (m2 & 1, (i, j)), (m2 & 2, (i, j2))):
                        wrong_variants += bool(said_disjoint) and truth[key]
    assert checked > 20000 and wrong == 0 and wrong_variants == 0, (checked, wrong, wrong_variants)
    assert caught >= 0.9 * apart, (caught, apart)             # (5 % and 50 % beyond touching along d: the filter is not vacuous)


def _host_filters(tmp_path):
    """tests/host_filters.cpp + the marked regions of csrc/collide_device.cuh, compiled for the CPU."""
    here = os.path.dirname(os.path.abspath(__file__))
    src = open(os.path.join(here, "..", "motionplanningtoolkit_b200", "csrc", "collide_device.cuh")).read()
    regions = re.findall(r"// \[host-test:begin\].*?\n(.*?)// \[host-test:end\]", src, flags=re.S)
    assert len(regions) == 8 and "exact_obb_disjoint" in regions[0] and "exact_tri_intersect_warp" in regions[1] and "exact_path_reaches" in regions[2]
    assert "tri_filter_f32" in regions[4] and "obb_disjoint_f32" in regions[4] and "rel_element" in regions[5] and "pose_world_box" in regions[6]
    assert "make_candidate_record2_rt" in regions[7]
    (tmp_path / "collide_filters_extracted.inc").write_text("\n".join(regions))
    cpp = tmp_path / "host_filters.cpp"
    cpp.write_text(open(os.path.join(here, "host_filters.cpp")).read())
    so = str(tmp_path / "host_filters.so")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-shared", "-fPIC",
                           "-I", os.path.join(here, "..", "motionplanningtoolkit_b200", "csrc"), str(cpp), "-o", so])
    L = ctypes.CDLL(so)
    f32p = ctypes.POINTER(ctypes.c_float)
    for fn in (L.host_obb_disjoint_f32, L.host_tri_filter_f32):
        fn.restype = ctypes.c_int
        fn.argtypes = [f32p, f32p, f32p]
    f64p = ctypes.POINTER(ctypes.c_double)
    L.host_decode_pose.restype = None
    L.host_decode_pose.argtypes = [f64p, ctypes.c_uint, f64p, f64p, f64p]
    L.host_exact_obb_disjoint.restype = ctypes.c_int
    L.host_exact_obb_disjoint.argtypes = [f64p, f64p, f64p]
    L.host_exact_tri_intersect.restype = ctypes.c_int
    L.host_exact_tri_intersect.argtypes = [f64p, f64p, f64p]
    u32p = ctypes.POINTER(ctypes.c_uint32)
    L.host_exact_path_reaches.restype = ctypes.c_int
    L.host_exact_path_reaches.argtypes = [f64p, f64p, u32p, ctypes.c_uint, f64p, u32p, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint]
    L.host_candidate_record.restype = None
    L.host_candidate_record.argtypes = [f64p, f64p, f64p, f64p, ctypes.c_uint, f32p]
    L.host_candidate_record2.restype = ctypes.c_int
    L.host_candidate_record2.argtypes = [f64p, f64p, f64p, f64p, ctypes.c_uint, f32p, f64p, f32p]
    L.host_pose_world_box.restype = None
    L.host_pose_world_box.argtypes = [f64p, ctypes.c_uint, f64p, f64p, f32p, f32p]
    L.host_relative_transform.restype = None
    L.host_relative_transform.argtypes = [f64p, f64p, f64p, f64p, f64p]
    return L


def test_device_object_box_and_relative_transform_equal_the_oracle_bit_for_bit(built, tmp_path):
    """pose_world_box (fcl::CollisionObject::computeAABB: translated tight box for an identity quaternion, sphere box around
    the transformed centre otherwise) and rel_element (fcl::relativeTransform) as the kernels evaluate them, compiled for
    the CPU from collide_device.cuh: the oracle's bits; the fp32 copy of the box the broadphase walks with contains it."""
    L = _host_filters(tmp_path)
    f64 = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    f32 = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    rng = np.random.default_rng(9001)
    for name, how in (("blimp", "merge"), ("blimp", "last"), ("unit_box", "last")):
        v, t = meshio.agent_submesh(meshio.load_mptm(name), how)
        om = O.Model(v, t)
        box, ctr, rad = om.local_aabb()
        model10 = np.concatenate([box, ctr, [rad]])
        for k in range(1500):
            T = rng.normal(size=3) * 10.0 ** rng.integers(-3, 4)
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            if k % 5 == 0:
                q = np.array([1.0, 0, 0, 0]) if k % 10 == 0 else np.array([-1.0, 0, 0, 0])
            for fmt_o, fmt_d, rec in ((O.POSE_TQ7, 1, np.concatenate([T, q])),
                                      (O.POSE_RT12, 0, np.concatenate([O.pose_decode(np.concatenate([T, q]), O.POSE_TQ7)[0].ravel(), T])),
                                      (O.POSE_RT12, 2, T.copy())):
                want = om.object_aabb(np.concatenate([np.eye(3).ravel(), T]) if fmt_d == 2 else rec, fmt_o)
                got, amin, amax = np.empty(6), np.empty(3, np.float32), np.empty(3, np.float32)
                L.host_pose_world_box(f64(rec), fmt_d, f64(model10), f64(got), f32(amin), f32(amax))
                assert got.tobytes() == want.tobytes(), (name, fmt_d, q)
                assert np.all(amin.astype(np.float64) <= got[:3]) and np.all(amax.astype(np.float64) >= got[3:])
                assert np.all(np.nextafter(amin, np.float32(np.inf)).astype(np.float64) > got[:3])
    for _ in range(3000):
        recs = []
        for _ in range(2):
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            T = rng.normal(size=3) * 10.0 ** rng.integers(-3, 4)
            R, T_, _q = O.pose_decode(np.concatenate([T, q]), O.POSE_TQ7)
            recs.append(np.concatenate([R.ravel(), T_]))
        Ro, To = O.relative_transform(recs[0], recs[1])
        out = np.empty(12)
        L.host_relative_transform(f64(recs[0][:9].copy()), f64(recs[0][9:].copy()), f64(recs[1][:9].copy()), f64(recs[1][9:].copy()), f64(out))
        assert out[:9].tobytes() == Ro.ravel().tobytes() and out[9:].tobytes() == To.tobytes()


def test_device_fp64_restatement_equals_the_oracle_bit_for_bit(built, tmp_path):
    """decode_pose (fcl::Transform3f from a matrix / a quaternion, Quaternion3f::fromRotation / toRotation) and
    exact_obb_disjoint (FCL's obbDisjoint with its operation order) as the kernels evaluate them, compiled for the CPU
    from collide_device.cuh: the same bits as the oracle for every pose, and the same decision on box pairs one bit
    before and one bit after touching."""
    L = _host_filters(tmp_path)
    f64 = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    rng = np.random.default_rng(424242)
    # ---- poses: random rotations, the quaternion branches (trace <= 0 with each largest diagonal element), +-identity,
    # unnormalised quaternions (FCL does not normalise), tiny and huge translations
    quats = [np.array(q, float) for q in ([1, 0, 0, 0], [-1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [0.5, 0.5, 0.5, 0.5],
                                           [2, 0, 0, 0], [1e-3, 1, 0, 0], [0.7, -0.1, 0.7, 0.1])]
    for _ in range(3000):
        q = rng.normal(size=4)
        quats.append(q / np.linalg.norm(q) * (1.0 if rng.random() < 0.8 else rng.uniform(0.5, 2.0)))
    for q in quats:
        T = rng.normal(size=3) * 10.0 ** rng.integers(-6, 6)
        tq7 = np.concatenate([T, q])
        Ro, To, qo = O.pose_decode(tq7, O.POSE_TQ7)
        R, T2, q2 = np.empty(9), np.empty(3), np.empty(4)
        L.host_decode_pose(f64(tq7), 1, f64(R), f64(T2), f64(q2))
        assert R.tobytes() == Ro.ravel().tobytes() and T2.tobytes() == To.tobytes(), q
        rt12 = np.concatenate([Ro.ravel(), T])          # the matrix form of the same pose: q = fromRotation(R)
        Ro2, To2, qo2 = O.pose_decode(rt12, O.POSE_RT12)
        L.host_decode_pose(f64(rt12), 0, f64(R), f64(T2), f64(q2))
        assert R.tobytes() == Ro2.ravel().tobytes() and q2.tobytes() == qo2.tobytes(), q
    # the planar formats are the matrix format with R(0,0)=c R(0,1)=s R(1,0)=-s R(1,1)=c
    for _ in range(200):
        th = rng.uniform(-7, 7)
        c, s_ = np.cos(th), np.sin(th)
        x = rng.normal(size=3)
        rt12 = np.array([c, s_, 0, -s_, c, 0, 0, 0, 1, *x])
        Ro, To, qo = O.pose_decode(rt12, O.POSE_RT12)
        R, T2, q2 = np.empty(9), np.empty(3), np.empty(4)
        L.host_decode_pose(f64(np.array([*x, c, s_])), 3, f64(R), f64(T2), f64(q2))
        assert R.tobytes() == Ro.ravel().tobytes() and T2.tobytes() == To.tobytes() and q2.tobytes() == qo.tobytes()
        L.host_decode_pose(f64(x.copy()), 2, f64(R), f64(T2), f64(q2))
        assert np.array_equal(R, np.eye(3).ravel()) and np.array_equal(q2, [1, 0, 0, 0]) and T2.tobytes() == x.tobytes()
    # ---- obbDisjoint at the last bit
    ma = _device_model(*meshio.agent_submesh(meshio.load_mptm("blimp"), "merge"))
    me = _device_model(*sorted(meshio.environment_submeshes(meshio.load_mptm("divider")), key=lambda m: -len(m[1]))[0])
    n64 = lambda nd, k: np.concatenate([nd["axis"][k].ravel(), nd["To"][k], nd["extent"][k], [float((nd["extent"][k] ** 2).sum())]])
    checked = 0
    for _ in range(600):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        R = scenes.quat_to_matrix(q[None])[0]
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        i, j = rng.integers(me["hb"].num_nodes), rng.integers(ma["hb"].num_nodes)
        bE = (me["nd"]["axis"][i], me["nd"]["To"][i], me["nd"]["extent"][i])
        bA = (ma["nd"]["axis"][j], ma["nd"]["To"][j], ma["nd"]["extent"][j])
        base = bE[1] - R @ bA[1]
        lo, hi = 0.0, 2.0 * (np.linalg.norm(bE[2]) + np.linalg.norm(bA[2])) + 1e-3
        if not O.obb_overlap(R, base, bE, bA) or O.obb_overlap(R, base + hi * d, bE, bA):
            continue
        for _ in range(70):
            mid = 0.5 * (lo + hi)
            lo, hi = (mid, hi) if O.obb_overlap(R, base + mid * d, bE, bA) else (lo, mid)
        for t in (0.0, lo * (1 - 1e-12), lo, hi, hi * (1 + 1e-12), 2 * hi):
            T = base + t * d
            rel = np.concatenate([R.ravel(), T])
            assert bool(L.host_exact_obb_disjoint(f64(rel), f64(n64(me["nd"], i)), f64(n64(ma["nd"], j)))) == (not O.obb_overlap(R, T, bE, bA))
            checked += 1
    assert checked > 2000
    # ---- Intersect::intersect_Triangle at the last bit (17 axes over 17 lanes, one ballot)
    checked = 0
    for _ in range(1500):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        R = scenes.quat_to_matrix(q[None])[0]
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        P = me["v"][me["t"][rng.integers(len(me["t"]))]]
        Q = ma["v"][ma["t"][rng.integers(len(ma["t"]))]]
        base = rng.dirichlet([1, 1, 1]) @ P - R @ (rng.dirichlet([1, 1, 1]) @ Q)
        lo, hi = 0.0, 2.0 * (np.ptp(P, 0).max() + np.ptp(Q, 0).max()) + 1e-3
        bisect = O.tri_intersect(P, Q, R, base) and not O.tri_intersect(P, Q, R, base + hi * d)
        for _ in range(70 if bisect else 0):
            mid = 0.5 * (lo + hi)
            lo, hi = (mid, hi) if O.tri_intersect(P, Q, R, base + mid * d) else (lo, mid)
        for t in ((0.0, lo * (1 - 1e-12), lo, hi, hi * (1 + 1e-12), 2 * hi) if bisect else (0.0, hi)):
            T = base + t * d
            rel = np.concatenate([R.ravel(), T])
            Pc, Qc = np.ascontiguousarray(P.ravel()), np.ascontiguousarray(Q.ravel())
            assert bool(L.host_exact_tri_intersect(f64(rel), f64(Pc), f64(Qc))) == O.tri_intersect(P, Q, R, T)
            checked += 1
    assert checked > 4000


def _round_up_f32(x):
    f = np.float32(x)
    return f if float(f) >= x else np.nextafter(f, np.float32(np.inf))


def _device_model(v, t):
    """What HostPack::add (csrc/mptcc_api.cpp) gives the device for one model: recentring origin, node32 and tri32
    records, r2 and rinf."""
    hb = capi.HostBvh(v, t)
    nd = hb.nodes()
    ctr = hb.local_aabb()[1]
    n = hb.num_nodes
    node32 = np.zeros((n, 4, 4), np.float32)
    node32[:, :3, :3] = nd["axis"]
    node32[:, :3, 3] = nd["To"] - ctr
    ext = nd["extent"] * (1.0 + 1e-6)
    e32 = ext.astype(np.float32)
    e32 = np.where(e32.astype(np.float64) >= ext, e32, np.nextafter(e32, np.float32(np.inf)))
    node32[:, 3, :3] = e32
    vv = np.asarray(v, np.float64)
    tri32 = np.zeros((len(t), 3, 4), np.float32)
    tri32[:, :, :3] = vv[np.asarray(t)] - ctr
    d = vv - ctr
    r2 = _round_up_f32(np.sqrt((d * d).sum(1).max()) * (1.0 + 1e-6))
    rinf = _round_up_f32(np.abs(d).max() * (1.0 + 1e-6))
    return dict(hb=hb, nd=nd, ctr=ctr, node32=node32, tri32=tri32, r2=r2, rinf=rinf, v=vv, t=np.asarray(t))


def _record13(R, T, me, ma):
    """R (9), Tc (3), S of the candidate record (make_candidate_record, collide_device.cuh)."""
    tc = R @ ma["ctr"] + T - me["ctr"]
    tf = tc.astype(np.float32)
    tinf = np.float32(np.abs(tf).max())
    S = np.float32(max(np.float32(5.01) * ma["r2"] + np.float32(4.0) * tinf * np.float32(1.0000002), me["rinf"])) + me["rinf"]
    return np.concatenate([R.ravel().astype(np.float32), tf, [S]]).astype(np.float32)


def test_fused_kernel_fp32_filters_are_one_sided_against_the_oracle(built, tmp_path):
    """obb_disjoint_f32 (64-byte fp32 nodes: fused kernel, round-1 narrowphase) and the three-valued triangle filter of
    every kernel, compiled for the CPU from collide_device.cuh itself, against FCL's fp64 tests as the oracle restates
    them, on pairs pushed to the last bit of touching: "disjoint" / "separated" only where FCL separates, "certain
    contact" only where FCL's Intersect::intersect_Triangle reports one; and both filters settle nearly every pair that
    is clearly apart or clearly crossing."""
    L = _host_filters(tmp_path)
    f32 = lambda a: np.ascontiguousarray(a, np.float32).ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    rng = np.random.default_rng(77031)
    ma = _device_model(*meshio.agent_submesh(meshio.load_mptm("blimp"), "merge"))
    envs = sorted(meshio.environment_submeshes(meshio.load_mptm("divider")), key=lambda m: -len(m[1]))[:2] + scenes.synthetic_environment(2)
    obb_checked = obb_wrong = obb_apart = obb_caught = 0
    tri_checked = tri_wrong = tri_clear = tri_settled = tri_deep = tri_certain = 0
    for ev, et in envs:
        me = _device_model(ev, et)
        ndE, ndA = me["nd"], ma["nd"]
        for _ in range(300):
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            R = scenes.quat_to_matrix(q[None])[0]
            d = rng.normal(size=3); d /= np.linalg.norm(d)
            # ---- boxes
            i, j = rng.integers(me["hb"].num_nodes), rng.integers(ma["hb"].num_nodes)
            bE = (ndE["axis"][i], ndE["To"][i], ndE["extent"][i])
            bA = (ndA["axis"][j], ndA["To"][j], ndA["extent"][j])
            base = ndE["To"][i] - R @ ndA["To"][j]
            lo, hi = 0.0, 2.0 * (np.linalg.norm(bE[2]) + np.linalg.norm(bA[2])) + 1e-3
            if O.obb_overlap(R, base, bE, bA) and not O.obb_overlap(R, base + hi * d, bE, bA):
                for _ in range(60):
                    mid = 0.5 * (lo + hi)
                    lo, hi = (mid, hi) if O.obb_overlap(R, base + mid * d, bE, bA) else (lo, mid)
                for t in (lo, lo * (1 - 1e-9), lo * (1 - 1e-6), lo * (1 - 1e-3), hi, hi * (1 + 1e-6), hi * (1 + 1e-3), hi * 1.05 + 1e-3, hi * 1.5 + 1e-2):
                    T = base + t * d
                    said = L.host_obb_disjoint_f32(f32(me["node32"][i]), f32(ma["node32"][j]), f32(_record13(R, T, me, ma)))
                    overlap = O.obb_overlap(R, T, bE, bA)
                    obb_checked += 1
                    obb_wrong += bool(said) and overlap
                    if t >= hi * 1.05:
                        obb_apart += 1
                        obb_caught += bool(said)
            # ---- triangles: a point of P and a point of Q coincide at t = 0, then Q moves away along d
            a, b = rng.integers(len(me["t"])), rng.integers(len(ma["t"]))
            P, Q = me["v"][me["t"][a]], ma["v"][ma["t"][b]]
            wp, wq = rng.dirichlet([1, 1, 1]), rng.dirichlet([1, 1, 1])
            base = wp @ P - R @ (wq @ Q)
            lo, hi = 0.0, 2.0 * (np.ptp(P, 0).max() + np.ptp(Q, 0).max()) + 1e-3
            if O.tri_intersect(P, Q, R, base) and not O.tri_intersect(P, Q, R, base + hi * d):
                for _ in range(60):
                    mid = 0.5 * (lo + hi)
                    lo, hi = (mid, hi) if O.tri_intersect(P, Q, R, base + mid * d) else (lo, mid)
                size = max(np.ptp(P, 0).max(), np.ptp(Q, 0).max())
                band = 1e-3 * float(_record13(R, base + hi * d, me, ma)[12])   # 100 x the filter's uncertainty in distance (~1e-5 S)
                for t, clear in ((lo, 0), (lo * (1 - 1e-9), 0), (lo * (1 - 1e-6), 0), (lo * (1 - 1e-3), 0), (hi, 0), (hi * (1 + 1e-6), 0),
                                 (hi * (1 + 1e-3), 0), (hi + 0.1 * size + band, 1), (0.5 * lo, 2), (hi + size + band, 1)):
                    T = base + t * d
                    v = L.host_tri_filter_f32(f32(me["tri32"][a]), f32(ma["tri32"][b]), f32(_record13(R, T, me, ma)))
                    truth = O.tri_intersect(P, Q, R, T)
                    tri_checked += 1
                    tri_wrong += (v == 0 and truth) or (v == 2 and not truth)
                    tri_clear += clear == 1
                    tri_settled += clear == 1 and v == 0
                    tri_deep += clear == 2 and truth
                    tri_certain += clear == 2 and truth and v == 2
    assert obb_checked > 5000 and obb_wrong == 0, (obb_checked, obb_wrong)
    assert obb_caught >= 0.9 * obb_apart, (obb_caught, obb_apart)
    assert tri_checked > 5000 and tri_wrong == 0, (tri_checked, tri_wrong)
    assert tri_settled >= 0.9 * tri_clear, (tri_settled, tri_clear)            # clearly apart: "separated"
    assert tri_certain >= 0.25 * tri_deep, (tri_certain, tri_deep)            # crossing half-way in: a good part "certain contact"


def test_device_path_replay_equals_the_oracles_traversal(built, tmp_path):
    """exact_path_reaches (the replay of FCL's root-to-leaf OBB tests for one triangle pair: two ancestor chains merged under
    firstOverSecond with warp shuffles, 32 fp64 tests per round) run on the CPU by replaying a converged warp lane by lane
    (tests/host_filters.cpp), against the oracle's own collisionRecurse with a probe on that leaf pair -- for triangle
    pairs that touch (the case the kernels replay), with the agent pushed out until the path breaks."""
    L = _host_filters(tmp_path)
    f64 = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    u32 = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
    rng = np.random.default_rng(86)
    av, at = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    ev, et = sorted(meshio.environment_submeshes(meshio.load_mptm("divider")), key=lambda m: -len(m[1]))[0]
    hbA, hbE = capi.HostBvh(av, at), capi.HostBvh(ev, et)
    n64A, ancA = hbA.exact_arrays(True)                     # the agent's device nodes are numbered breadth-first
    n64E, ancE = hbE.exact_arrays(False)
    oA, oE = O.Model(av, at), O.Model(ev, et)
    av, ev, at, et = np.asarray(av, float), np.asarray(ev, float), np.asarray(at), np.asarray(et)
    reached = broken = 0
    for _ in range(120):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        R = scenes.quat_to_matrix(q[None])[0]
        p1, p2 = int(rng.integers(len(et))), int(rng.integers(len(at)))
        P, Q = ev[et[p1]], av[at[p2]]
        base = rng.dirichlet([1, 1, 1]) @ P - R @ (rng.dirichlet([1, 1, 1]) @ Q)      # the two triangles share a point
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        for t in (0.0, 1e-3, 0.05, 0.3, 1.0, 3.0):
            T = base + t * d
            rel = np.concatenate([R.ravel(), T])
            got = L.host_exact_path_reaches(f64(rel), f64(n64E), u32(ancE), ancE.shape[1], f64(n64A), u32(ancA), ancA.shape[1], p1, p2)
            want = O.path_reaches(oE, oA, R, T, p1, p2)
            assert got == int(want), (p1, p2, t, got, want)
            reached += want
            broken += not want
    assert reached > 100 and broken > 100, (reached, broken)


def test_path_certificate_implies_fcl_reaches_the_pair(built, tmp_path):
    """DESIGN.md section 7 on the CPU: where the kernels skip the path replay of a contact -- min(mu_chain) of the two
    triangles >= the candidate's threshold, both as the library computes them -- FCL's traversal (the oracle's, with a probe on
    the leaf pair) does reach that pair.  Checked on triangle pairs of the fixtures that share a point (most of them are
    certified: the test is not vacuous)."""
    L = _host_filters(tmp_path)
    f64 = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    f32 = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    rng = np.random.default_rng(4711)
    av, at = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    certified = contacts = 0
    for ev, et in sorted(meshio.environment_submeshes(meshio.load_mptm("divider")), key=lambda m: -len(m[1]))[:2] + scenes.synthetic_environment(1):
        hbA, hbE = capi.HostBvh(av, at), capi.HostBvh(ev, et)
        muA, cA = hbA.certificate_data()
        muE, cE = hbE.certificate_data()
        oA, oE = O.Model(av, at), O.Model(ev, et)
        va, ve, ta, te = np.asarray(av, float), np.asarray(ev, float), np.asarray(at), np.asarray(et)
        m1v = np.array([*cE["ctr"], 1.0, cE["cmax1"], cE["esum"]]); m2v = np.array([*cA["ctr"], cA["r2"], cA["cmax1"], cA["esum"]])
        obj13 = np.array([*np.eye(3).ravel(), 0, 0, 0, 1.0])
        for k in range(250):
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            R = scenes.quat_to_matrix(q[None])[0]
            p1, p2 = int(rng.integers(len(te))), int(rng.integers(len(ta)))
            P, Q = ve[te[p1]], va[ta[p2]]
            T = rng.dirichlet([1, 1, 1]) @ P - R @ (rng.dirichlet([1, 1, 1]) @ Q)
            if not O.tri_intersect(P, Q, R, T):
                continue
            contacts += 1
            rec, rel, cert = np.empty(16, np.float32), np.empty(12), np.empty(1, np.float32)
            pose = np.concatenate([R.ravel(), T])
            assert L.host_candidate_record2(f64(obj13), f64(m1v), f64(m2v), f64(pose), 0, f32(rec), f64(rel), f32(cert)) == 1
            if min(muE[p1], muA[p2]) >= cert[0]:
                certified += 1
                assert O.path_reaches(oE, oA, R, T, p1, p2), (p1, p2)
    assert contacts > 400 and certified > 0.8 * contacts, (contacts, certified)


def test_candidate_record_of_the_broadphase_kernel(built, tmp_path):
    """make_candidate_record2 (collide_device.cuh) compiled for the CPU: the fp64 relative transform it passes on is the
    oracle's relativeTransform bit for bit (for an object at the identity: the pose itself), the fp32 record is that
    transform in the recentred frames rounded once (R, Tc) with the inverse translation, the triangle-filter scale and the
    path-certificate threshold on the safe side, and poses that are not rigid are flagged."""
    L = _host_filters(tmp_path)
    f64 = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    f32 = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    rng = np.random.default_rng(5150)
    ulp = lambda x: np.abs(np.nextafter(np.float32(x), np.float32(np.inf)) - np.float32(x)).astype(np.float64)
    for k in range(4000):
        ident = k % 2 == 0
        q1 = np.array([1.0, 0, 0, 0]) if ident else rng.normal(size=4)
        q1 /= np.linalg.norm(q1)
        T1 = np.zeros(3) if ident else rng.normal(size=3) * 5
        q2 = rng.normal(size=4); q2 /= np.linalg.norm(q2)
        T2 = rng.normal(size=3) * 10.0 ** rng.integers(-2, 3)
        R1, _, _ = O.pose_decode(np.concatenate([T1, q1]), O.POSE_TQ7)
        c1, c2 = rng.normal(size=3) * 3, rng.normal(size=3)
        rinf, r2 = abs(rng.normal()) * 2 + 0.1, abs(rng.normal()) * 2 + 0.1
        m1v = np.array([*c1, np.float32(rinf), 3.0, 2.0]); m2v = np.array([*c2, np.float32(r2), 1.5, 0.7])
        obj13 = np.array([*R1.ravel(), *T1, 1.0 if ident else 0.0])
        pose = np.concatenate([T2, q2])
        rec, rel, cert = np.empty(16, np.float32), np.empty(12), np.empty(1, np.float32)
        rigid = L.host_candidate_record2(f64(obj13), f64(m1v), f64(m2v), f64(pose), 1, f32(rec), f64(rel), f32(cert))
        Ro, To = O.relative_transform(np.concatenate([T1, q1]), pose, O.POSE_TQ7, O.POSE_TQ7)
        assert rigid == 1 and np.array_equal(rel[:9], Ro.ravel()) and np.array_equal(rel[9:], To)
        assert np.array_equal(rec[:9], Ro.ravel().astype(np.float32))
        tc = Ro @ c2 + To - c1
        assert np.all(np.abs(rec[9:12].astype(np.float64) - tc) <= ulp(tc) + 1e-15 * np.abs(tc).max())
        ti = -(Ro.T @ tc)
        assert np.all(np.abs(rec[12:15].astype(np.float64) - ti) <= ulp(ti) + 4e-15 * np.abs(tc).sum())
        tinf = np.abs(rec[9:12]).max()
        assert rec[15] >= max(5.01 * np.float32(r2) + 4.0 * tinf, np.float32(rinf)) + np.float32(rinf) * (1 - 1e-6)
        mag = np.abs(tc).sum() + np.abs(c1).sum() + 1.75 * np.abs(c2).sum() + 3.0 + 1.5 + 2.0 + 0.7
        assert cert[0] >= 1e-5 * mag
        # the records the filter tests above build in numpy are this record
        r16 = np.concatenate([Ro.ravel(), tc, ti, [0.0]]).astype(np.float32)
        assert np.all(np.abs(r16[:15].astype(np.float64) - rec[:15].astype(np.float64)) <= ulp(r16[:15]) + 1e-12)
    # the record of the fused kernel, as the filter tests above build it in numpy (_record13)
    ma = _device_model(*meshio.agent_submesh(meshio.load_mptm("blimp"), "merge"))
    me = _device_model(*meshio.environment_submeshes(meshio.load_mptm("divider"))[3])
    for _ in range(500):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        R = scenes.quat_to_matrix(q[None])[0]
        T = rng.normal(size=3) * 10.0 ** rng.integers(-2, 3)
        rec13 = np.empty(13, np.float32)
        L.host_candidate_record(f64(np.array([*np.eye(3).ravel(), 0, 0, 0, 1.0])), f64(np.array([*me["ctr"], me["rinf"], 0, 0])),
                                f64(np.array([*ma["ctr"], ma["r2"], 0, 0])), f64(np.concatenate([R.ravel(), T])), 0, f32(rec13))
        want = _record13(R, T, me, ma)
        assert np.array_equal(rec13[:9], want[:9]) and np.all(np.abs(rec13[9:12].astype(np.float64) - want[9:12]) <= ulp(want[9:12]))
        assert abs(float(rec13[12]) - float(want[12])) <= 4 * ulp(want[12])
    # not rigid: a scaled, a sheared and a NaN rotation block (RT12 records)
    for bad in (np.eye(3) * (1 + 1e-9), np.array([[1, 1e-8, 0], [0, 1, 0], [0, 0, 1.0]]), np.full((3, 3), np.nan)):
        pose = np.concatenate([bad.ravel(), [0.1, 0.2, 0.3]])
        rec, rel, cert = np.empty(16, np.float32), np.empty(12), np.empty(1, np.float32)
        obj13 = np.array([*np.eye(3).ravel(), 0, 0, 0, 1.0])
        assert L.host_candidate_record2(f64(obj13), f64(np.array([0, 0, 0, 1.0, 1, 1])), f64(np.array([0, 0, 0, 1.0, 1, 1])), f64(pose), 0,
                                        f32(rec), f64(rel), f32(cert)) == 0


def test_broadphase_grid_and_tree_list_every_overlapping_object(built):
    """The candidate set: for any pose, the objects whose world AABB overlaps the agent's (FCL's closed-interval
    AABB::overlap, what the dynamic AABB tree reports) must all be in what the broadphase structures list for the pose box --
    the grid cell of its minimum corner, and the clusters the threaded tree walk reaches.  Built on the host by a context
    without devices, looked up the way the kernels do, compared with the oracle's object boxes."""
    rng = np.random.default_rng(31337)
    agent = meshio.agent_submesh(meshio.load_mptm("blimp"), "merge")
    om = O.Model(*agent)
    for env, rot in ((scenes.synthetic_environment(160), None), (meshio.environment_submeshes(meshio.load_mptm("divider")), None),
                     (meshio.environment_submeshes(meshio.load_mptm("divider")), (0.3, -0.2, 0.1, 0.9, 0.1, -0.3, 0.2))):
        pose7 = (0, 0, 0, 1, 0, 0, 0) if rot is None else tuple(np.concatenate([rot[:3], np.array(rot[3:]) / np.linalg.norm(rot[3:])]))
        hs = capi.HostScene()
        hs.set_environment(env, pose7)
        hs.set_agent([agent])
        with pytest.raises(capi.MptccError):               # no device: nothing can be checked with it
            hs.check_poses(np.zeros((1, 12)))
        oenv = O.Environment([O.Model(v, t) for v, t in env], pose7)
        boxes = np.array([oenv.object_aabb(i) for i in range(len(env))])
        lo, hi = boxes[:, :3].min(0), boxes[:, 3:].max(0)
        n_pairs = n_grid_listed = 0
        for k in range(1500):
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            if k % 7 == 0:
                q = np.array([1.0, 0, 0, 0])               # identity quaternion: the tight box instead of the sphere box
            T = rng.uniform(lo - 2.0, hi + 2.0)
            if k % 3 == 0:                                  # ... and boxes that just touch an object's box
                b = boxes[rng.integers(len(boxes))]
                T = b[rng.integers(2) * 3:][:3] + rng.normal(size=3) * 0.5
            pb = om.object_aabb(np.concatenate([T, q]), O.POSE_TQ7)
            want = {tuple(b) for b in boxes if np.all(b[:3] <= pb[3:]) and np.all(b[3:] >= pb[:3])}
            n_pairs += len(want)
            ng, got_g, listed_g = hs.grid_lookup(pb)
            nt, got_t, listed_t = hs.tree_lookup(pb)
            assert nt >= 0 and {tuple(b) for b in got_t} == want, (k, nt, len(want))
            assert ng == -2 or (ng >= 0 and {tuple(b) for b in got_g} == want), (k, ng, len(want))
            assert ng != -2                                  # finite poses inside the scene are always grid-addressable
            n_grid_listed += listed_g
        assert n_pairs > 500 and n_grid_listed >= n_pairs
        # an exactly touching box (closed intervals) is a candidate; a box one ulp away is not
        b = boxes[0]
        touch = np.concatenate([[b[3], b[1], b[2]], [b[3] + 0.5, b[4], b[5]]])
        away = touch.copy(); away[0] = np.nextafter(b[3], np.inf)
        assert tuple(b) in {tuple(x) for x in hs.tree_lookup(touch)[1]} and tuple(b) not in {tuple(x) for x in hs.tree_lookup(away)[1]}
        g = hs.grid_lookup(touch)
        assert g[0] == -2 or tuple(b) in {tuple(x) for x in g[1]}
        hs.close()


def test_empty_mesh_is_rejected(built):
    with pytest.raises(capi.MptccError):
        capi.HostBvh(np.zeros((0, 3)), np.zeros((0, 3), np.uint32))


def test_agent_submesh_rules():
    subs = meshio.load_mptm("blimp")
    assert [len(t) for _, t in subs] == [8, 121, 1194, 32]
    v, t = meshio.agent_submesh(subs, "last")            # the reference keeps only the last sub-mesh (Q2)
    assert len(t) == 32
    v, t = meshio.agent_submesh(subs, "merge")
    assert len(t) == 1355 and t.max() == len(v) - 1
    assert len(meshio.environment_submeshes(meshio.load_mptm("divider"))) == 110


def test_omnidirectional_edge_interpolation():
    """Omnidirectional::getPoses: iterations = (unsigned)(dist/dt); end point added iff iterations*dt < dist."""
    # dist/dt = 10 exactly: ten poses 0 .. 0.9 and NO end pose, because 10*0.1 < 1.0 is false (reference quirk)
    p = scenes.omnidirectional_edge_poses([0, 0, 0], [1, 0, 0], 0.1)
    assert len(p) == 10 and np.allclose(p[0], 0) and np.allclose(p[-1], [0.9, 0, 0])
    p = scenes.omnidirectional_edge_poses([0, 0, 0], [0.05, 0, 0], 0.1)      # shorter than dt: start and end
    assert len(p) == 2
    p = scenes.omnidirectional_edge_poses([0, 0, 0], [0, 0.35, 0], 0.1)
    assert len(p) == 4 and np.allclose(p[:, 1], [0, 0.1, 0.2, 0.35])


def test_splitmix64_reference_values():
    r = scenes.SplitMix64(0)
    assert [int(x) for x in r.next_u64(3)] == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]
    a = scenes.SplitMix64(5).next_u64(10)
    r = scenes.SplitMix64(5)
    b = np.concatenate([r.next_u64(4), r.next_u64(6)])
    assert np.array_equal(a, b)


def test_shard_ranges_are_word_aligned_and_cover():
    for n in (0, 1, 31, 32, 33, 1000, 4097, 1 << 20):
        for w in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(w):
                b, e = sharding.shard_range(n, w, r)
                assert b == prev and b % 32 == 0 or b == n
                assert e >= b
                prev = e
            assert prev == n


def _gloo_worker(rank, world, port, n, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    box = meshio.load_mptm("unit_box")[0]
    om = O.Model(*box)
    env = O.Environment([om])
    rng = scenes.SplitMix64(1234)
    poses = scenes.random_poses_rt12(rng, n, (-1.6,) * 3, (1.6,) * 3)            # same batch on every rank
    local = sharding.check_sharded(lambda p: O.verdicts(env, [om], p), poses, world, rank)
    words = sharding.all_gather_verdicts(local, n, world, rank, dist)
    whole = O.verdicts(env, [om], poses)
    np.save(os.path.join(out_dir, "ok_%d.npy" % rank), np.array([np.array_equal(words, whole)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1000, 4097])
def test_two_rank_sharding_matches_unsplit_run(built, tmp_path, n):
    """N > 1 path on CPU (gloo, world_size 2): per-rank slices concatenate to the unsplit verdict words."""
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + n) % 2000
    mp.spawn(_gloo_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert np.load(os.path.join(str(tmp_path), "ok_%d.npy" % r))[0]


def test_cpp_mesh_loaders_match_the_python_converter(built, tmp_path):
    """include/mpt_b200/mesh_loaders.hpp (.dae / .3ds readers with Assimp's flags-0 sub-mesh rules)
    reproduces the committed .mptm fixtures byte for byte.  Needs the reference's mesh files (build container)."""
    ref = "/root/reference/mesh_models"
    if not os.path.isdir(ref):
        pytest.skip("reference fixtures are not present on this box")
    exe = str(tmp_path / "mesh_convert")
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "mesh_convert.cpp"), "-o", exe])
    for src, name in (("agent_models/unit_box.dae", "unit_box"), ("agent_models/blimp.3ds", "blimp"),
                      ("environment_models/model.dae", "model"), ("environment_models/divider.dae", "divider")):
        out = str(tmp_path / (name + ".mptm"))
        subprocess.check_call([exe, os.path.join(ref, src), out], stdout=subprocess.DEVNULL)
        assert open(out, "rb").read() == open(os.path.join(meshio.DATA_DIR, name + ".mptm"), "rb").read(), name
    # the flat format round-trips
    out = str(tmp_path / "rt.mptm")
    subprocess.check_call([exe, os.path.join(meshio.DATA_DIR, "blimp.mptm"), out], stdout=subprocess.DEVNULL)
    assert open(out, "rb").read() == open(os.path.join(meshio.DATA_DIR, "blimp.mptm"), "rb").read()


@pytest.mark.parametrize("example", ["omnidirectional_demo", "prm_roadmap", "self_collision_check", "mesh_convert",
                                     "grid_prmlite_demo", "rrt_speculative"])
def test_examples_compile_as_cxx11_against_the_shim(built, example, tmp_path):
    """The drop-in headers are C++11 like the reference (CMakeLists.txt: -std=c++11) and link against libmptcc only."""
    libdir = os.path.join(ROOT, "motionplanningtoolkit_b200")
    exe = str(tmp_path / example)
    r = subprocess.run(["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", example + ".cpp"), "-o", exe, "-L", libdir, "-lmptcc",
                        "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


def test_mesh_loaders_reject_malformed_files_and_read_polylists(built, tmp_path):
    """Mesh files are untrusted input: out-of-range face indices, chunks / counts that run past the
    file, a zero accessor stride must fail cleanly; <polylist> / <polygons> contribute their three-vertex faces."""
    import struct
    exe = str(tmp_path / "mesh_convert")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "mesh_convert.cpp"), "-o", exe])
    out = str(tmp_path / "o.mptm")

    def run(name, data):
        p = str(tmp_path / name)
        open(p, "wb").write(data)
        return subprocess.run([exe, p, out], capture_output=True, text=True, timeout=20)

    good = open(os.path.join(meshio.DATA_DIR, "unit_box.mptm"), "rb").read()
    assert run("ok.mptm", good).returncode == 0
    assert run("huge_counts.mptm", good[:12] + struct.pack("<II", 0x7fffffff, 0x7fffffff) + good[20:]).returncode != 0
    assert run("truncated.mptm", good[:len(good) // 2]).returncode != 0
    bad_idx = bytearray(good)
    bad_idx[-4:] = struct.pack("<I", 1000)                      # last triangle index beyond the 24+ vertices
    assert run("bad_index.mptm", bytes(bad_idx)).returncode != 0
    # 3DS: a vertex-list chunk (0x4110) that claims more vertices than the chunk holds
    verts = struct.pack("<HI", 0x4110, 6 + 2 + 12) + struct.pack("<H", 5000) + struct.pack("<3f", 0, 0, 0)
    tri = struct.pack("<HI", 0x4100, 6 + len(verts)) + verts
    obj = struct.pack("<HI", 0x4000, 6 + 2 + len(tri)) + b"a\0" + tri
    edit = struct.pack("<HI", 0x3D3D, 6 + len(obj)) + obj
    assert run("overrun.3ds", struct.pack("<HI", 0x4D4D, 6 + len(edit)) + edit).returncode != 0

    def dae(prim, stride="3"):
        return ("""<?xml version="1.0"?><COLLADA><library_geometries><geometry id="g"><mesh>
<source id="p"><float_array id="pa" count="15">0 0 0  1 0 0  1 1 0  0 1 0  0 0 1</float_array>
<technique_common><accessor source="#pa" count="5" stride="%s"/></technique_common></source>
<vertices id="v"><input semantic="POSITION" source="#p"/></vertices>%s</mesh></geometry></library_geometries>
<library_visual_scenes><visual_scene id="s"><node id="n"><instance_geometry url="#g"/></node></visual_scene>
</library_visual_scenes></COLLADA>""" % (stride, prim)).encode()

    r = run("stride0.dae", dae('<triangles count="1"><input semantic="VERTEX" source="#v" offset="0"/><p>0 1 2</p></triangles>', "0"))
    assert r.returncode != 0
    assert run("oob.dae", dae('<triangles count="1"><input semantic="VERTEX" source="#v" offset="0"/><p>0 1 9</p></triangles>')).returncode != 0
    r = run("plist.dae", dae('<polylist count="3"><input semantic="VERTEX" source="#v" offset="0"/><vcount>3 4 3</vcount>'
                             '<p>0 1 2  0 1 2 3  2 3 4</p></polylist>'))
    assert r.returncode == 0 and "1 sub-meshes, 2 triangles" in r.stdout, r.stdout + r.stderr
    r = run("pgons.dae", dae('<polygons count="2"><input semantic="VERTEX" source="#v" offset="0"/><p>0 1 2</p><p>0 1 2 3</p></polygons>'))
    assert r.returncode == 0 and "1 sub-meshes, 1 triangles" in r.stdout, r.stdout + r.stderr
    # the C ABI refuses a mesh whose triangle indices point past its vertices
    v, t = meshio.load_mptm("unit_box")[0]
    t = t.copy()
    t[0, 0] = len(v) + 5
    with pytest.raises(capi.MptccError):
        capi.HostBvh(v, t)


REFERENCE = "/root/reference"
PATCHES = sorted(f for f in os.listdir(os.path.join(ROOT, "patches")) if f.endswith(".diff"))


@pytest.mark.parametrize("patch_file", PATCHES)
def test_call_site_patches_apply_to_the_reference(patch_file):
    """the batched call sites are committed as patches against the reference's planners and
    discretizations (patches/*.diff); they must apply cleanly to the checkout they were made from."""
    if not os.path.isdir(REFERENCE):
        pytest.skip("the reference checkout is not present on this box")
    r = subprocess.run(["patch", "--dry-run", "-p1", "-d", REFERENCE, "-i", os.path.join(ROOT, "patches", patch_file)],
                       capture_output=True, text=True)
    assert r.returncode == 0 and "FAILED" not in r.stdout, r.stdout + r.stderr


def test_reference_grid_discretization_compiles_unchanged_and_patched(built, tmp_path):
    """The reference's discretizations/workspace/griddiscretization.hpp, UNMODIFIED, instantiated with Map3D from
    include/mpt_b200 (its constructor calls safePoses(agent, vector<vector<Transform3f>>), map3d.hpp:39-42), and again
    with patches/discretizations_workspace_griddiscretization_hpp.diff applied (one safePosesBatch for the grid).
    The binaries are left under examples/bin/ (git-ignored; they travel with the snapshot) for the GPU test, because
    the reference checkout does not exist on the GPU box."""
    if not os.path.isdir(REFERENCE):
        pytest.skip("the reference checkout is not present on this box")
    libdir = os.path.join(ROOT, "motionplanningtoolkit_b200")
    bindir = os.path.join(ROOT, "examples", "bin")
    os.makedirs(bindir, exist_ok=True)
    patched = tmp_path / "patched" / "discretizations" / "workspace"
    patched.mkdir(parents=True)
    src = os.path.join(REFERENCE, "discretizations", "workspace", "griddiscretization.hpp")
    open(str(patched / "griddiscretization.hpp"), "w").write(open(src).read())
    subprocess.check_call(["patch", "-s", "-p1", "-d", str(tmp_path / "patched"), "-i",
                           os.path.join(ROOT, "patches", "discretizations_workspace_griddiscretization_hpp.diff")])
    for name, inc in (("grid_prmlite_demo_reference", REFERENCE), ("grid_prmlite_demo_patched", str(tmp_path / "patched"))):
        r = subprocess.run(["g++", "-std=c++11", "-O1", "-DMPT_HAVE_REFERENCE", "-I", os.path.join(ROOT, "include"), "-I", inc,
                            os.path.join(ROOT, "examples", "grid_prmlite_demo.cpp"), "-o", os.path.join(bindir, name),
                            "-L", libdir, "-lmptcc", "-Wl,-rpath,$ORIGIN/../../motionplanningtoolkit_b200"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
