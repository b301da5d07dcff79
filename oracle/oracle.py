"""ctypes binding of the CPU oracle (oracle/liborc.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (motionplanningtoolkit_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

POSE_RT12 = 0
POSE_TQ7 = 1


class Counters(C.Structure):
    _fields_ = [("n_pose_sets", C.c_uint64), ("n_aabb_tests", C.c_uint64), ("n_obj_pairs", C.c_uint64),
                ("n_bv_tests", C.c_uint64), ("n_leaf_tests", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "fcl_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liborc.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, dp, u32p, i32p = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
        L.orc_model_build.restype = vp
        L.orc_model_build.argtypes = [dp, C.c_uint32, u32p, C.c_uint32]
        L.orc_model_free.argtypes = [vp]
        L.orc_model_num_nodes.restype = C.c_uint32
        L.orc_model_num_nodes.argtypes = [vp]
        L.orc_model_num_tris.restype = C.c_uint32
        L.orc_model_num_tris.argtypes = [vp]
        L.orc_model_get_nodes.argtypes = [vp, dp, dp, dp, i32p, i32p, i32p]
        L.orc_model_get_primitive_indices.argtypes = [vp, u32p]
        L.orc_model_get_local_aabb.argtypes = [vp, dp, dp, dp]
        L.orc_env_create.restype = vp
        L.orc_env_create.argtypes = [C.c_int, C.POINTER(vp), dp]
        L.orc_env_free.argtypes = [vp]
        L.orc_env_get_object_aabb.argtypes = [vp, C.c_int, dp]
        L.orc_is_in_collision.restype = C.c_int
        L.orc_is_in_collision.argtypes = [vp, C.c_int, C.POINTER(vp), C.c_uint64, dp, C.c_int, C.POINTER(Counters)]
        L.orc_verdicts.restype = C.c_int
        L.orc_verdicts.argtypes = [vp, C.c_int, C.POINTER(vp), C.c_uint64, dp, C.c_int, u32p, C.c_int,
                                   C.POINTER(Counters)]
        L.orc_verdicts_self.restype = C.c_int
        L.orc_verdicts_self.argtypes = L.orc_verdicts.argtypes
        L.orc_self_collision.restype = C.c_int
        L.orc_self_collision.argtypes = [C.c_int, C.POINTER(vp), dp, C.c_int, C.POINTER(Counters)]
        L.orc_verdicts_bruteforce.restype = C.c_int
        L.orc_verdicts_bruteforce.argtypes = [vp, C.c_int, C.POINTER(vp), C.c_uint64, dp, C.c_int, u32p]
        L.orc_verdicts_bruteforce_fast.restype = C.c_int
        L.orc_verdicts_bruteforce_fast.argtypes = [vp, C.c_int, C.POINTER(vp), C.c_uint64, dp, C.c_int, u32p, C.c_int]
        L.orc_obb_overlap.restype = C.c_int
        L.orc_obb_overlap.argtypes = [dp] * 8
        L.orc_tri_intersect.restype = C.c_int
        L.orc_tri_intersect.argtypes = [dp] * 4
        L.orc_pose_decode.argtypes = [dp, C.c_int, dp, dp, dp]
        L.orc_path_reaches.restype = C.c_int
        L.orc_path_reaches.argtypes = [vp, vp, dp, dp, C.c_int, C.c_int]
        L.orc_relative_transform.restype = None
        L.orc_relative_transform.argtypes = [dp, C.c_int, dp, C.c_int, dp, dp]
        L.orc_object_aabb.argtypes = [vp, dp, C.c_int, dp]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Model:
    """fcl::BVHModel<fcl::OBBRSS> built from one (vertices, triangles) sub-model."""

    def __init__(self, xyz, tri):
        self.xyz = _f64(xyz).reshape(-1, 3)
        self.tri = np.ascontiguousarray(tri, dtype=np.uint32).reshape(-1, 3)
        self.h = lib().orc_model_build(_dp(self.xyz), len(self.xyz), _u32p(self.tri), len(self.tri))
        if not self.h:
            raise ValueError("empty model")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_model_free(self.h)
            self.h = None

    @property
    def num_nodes(self):
        return lib().orc_model_num_nodes(self.h)

    def nodes(self):
        n = self.num_nodes
        axis = np.empty((n, 3, 3)); To = np.empty((n, 3)); ext = np.empty((n, 3))
        fc = np.empty(n, np.int32); fp = np.empty(n, np.int32); npr = np.empty(n, np.int32)
        i32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        lib().orc_model_get_nodes(self.h, _dp(axis), _dp(To), _dp(ext), i32(fc), i32(fp), i32(npr))
        return dict(axis=axis, To=To, extent=ext, first_child=fc, first_primitive=fp, num_primitives=npr)

    def primitive_indices(self):
        out = np.empty(len(self.tri), np.uint32)
        lib().orc_model_get_primitive_indices(self.h, _u32p(out))
        return out

    def local_aabb(self):
        box = np.empty(6); ctr = np.empty(3); r = C.c_double()
        lib().orc_model_get_local_aabb(self.h, _dp(box), _dp(ctr), C.byref(r))
        return box, ctr, r.value

    def object_aabb(self, rec, fmt=POSE_RT12):
        rec = _f64(rec)
        out = np.empty(6)
        lib().orc_object_aabb(self.h, _dp(rec), fmt, _dp(out))
        return out


class Environment:
    """StaticEnvironmentMeshHandler: objects = list of Model, pose7 = x y z qw qx qy qz (one for all
    objects, as the reference does, or one per object)."""

    def __init__(self, models, pose7=(0, 0, 0, 1, 0, 0, 0)):
        self.models = list(models)
        p = _f64(pose7)
        if p.ndim == 1:
            p = np.tile(p, (len(self.models), 1))
        self.pose7 = _f64(p)
        arr = (C.c_void_p * max(1, len(self.models)))(*[m.h for m in self.models])
        self.h = lib().orc_env_create(len(self.models), arr, _dp(self.pose7))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_env_free(self.h)
            self.h = None

    def object_aabb(self, i):
        out = np.empty(6)
        lib().orc_env_get_object_aabb(self.h, i, _dp(out))
        return out


def _links(links):
    return (C.c_void_p * len(links))(*[m.h for m in links])


def _poses(poses, n_links, fmt):
    rl = 7 if fmt == POSE_TQ7 else 12
    p = _f64(poses).reshape(-1, n_links, rl)
    return p, p.shape[0]


def is_in_collision(env, links, poses, fmt=POSE_RT12, counters=None):
    """MeshHandler::isInCollision: one bool for the whole batch of pose-sets."""
    p, n = _poses(poses, len(links), fmt)
    c = counters if counters is not None else Counters()
    return bool(lib().orc_is_in_collision(env.h, len(links), _links(links), n, _dp(p), fmt, C.byref(c)))


def verdicts(env, links, poses, fmt=POSE_RT12, n_threads=1, counters=None):
    """One isInCollision per pose-set; returns uint32 bit words (bit=1: in collision)."""
    p, n = _poses(poses, len(links), fmt)
    bits = np.zeros((n + 31) // 32, np.uint32)
    c = counters if counters is not None else Counters()
    lib().orc_verdicts(env.h, len(links), _links(links), n, _dp(p), fmt, _u32p(bits), n_threads, C.byref(c))
    return bits


def verdicts_self(env, links, poses, fmt=POSE_RT12, n_threads=1, counters=None):
    """verdicts() with checkSelfCollision = true: bit = self-collision of the pose-set OR environment collision."""
    p, n = _poses(poses, len(links), fmt)
    bits = np.zeros((n + 31) // 32, np.uint32)
    c = counters if counters is not None else Counters()
    lib().orc_verdicts_self(env.h, len(links), _links(links), n, _dp(p), fmt, _u32p(bits), n_threads, C.byref(c))
    return bits


def self_collision(links, pose_set, fmt=POSE_RT12, counters=None):
    """Do the links of ONE pose-set collide with each other (meshhandler.hpp:205-219)?"""
    p, n = _poses(pose_set, len(links), fmt)
    assert n == 1
    c = counters if counters is not None else Counters()
    return bool(lib().orc_self_collision(len(links), _links(links), _dp(p), fmt, C.byref(c)))


def verdicts_bruteforce(env, links, poses, fmt=POSE_RT12):
    p, n = _poses(poses, len(links), fmt)
    bits = np.zeros((n + 31) // 32, np.uint32)
    lib().orc_verdicts_bruteforce(env.h, len(links), _links(links), n, _dp(p), fmt, _u32p(bits))
    return bits


def verdicts_bruteforce_fast(env, links, poses, fmt=POSE_RT12, n_threads=8):
    """Structure-free oracle with a triangle-box prefilter and threads (see fcl_oracle.h)."""
    p, n = _poses(poses, len(links), fmt)
    bits = np.zeros((n + 31) // 32, np.uint32)
    lib().orc_verdicts_bruteforce_fast(env.h, len(links), _links(links), n, _dp(p), fmt, _u32p(bits), n_threads)
    return bits


def unpack_bits(bits, n):
    return ((np.repeat(bits, 32) >> np.tile(np.arange(32, dtype=np.uint32), len(bits))) & 1).astype(bool)[:n]


def obb_overlap(R0, T0, b1, b2):
    a = [_f64(x) for x in (R0, T0, b1[0], b1[1], b1[2], b2[0], b2[1], b2[2])]
    return bool(lib().orc_obb_overlap(*[_dp(x) for x in a]))


def tri_intersect(P, Q, R=np.eye(3), T=np.zeros(3)):
    a = [_f64(x) for x in (P, Q, R, T)]
    return bool(lib().orc_tri_intersect(*[_dp(x) for x in a]))


def path_reaches(m1, m2, R, T, prim1, prim2):
    """Does FCL's collisionRecurse on (m1, m2) under x1 = R x2 + T reach the leaf pair of triangles (prim1, prim2)?"""
    return bool(lib().orc_path_reaches(m1.h, m2.h, _dp(_f64(R)), _dp(_f64(T)), int(prim1), int(prim2)))


def relative_transform(rec1, rec2, fmt1=POSE_RT12, fmt2=POSE_RT12):
    """fcl::relativeTransform(tf1, tf2) of two pose records -> (R, T)."""
    R = np.empty((3, 3)); T = np.empty(3)
    lib().orc_relative_transform(_dp(_f64(rec1)), fmt1, _dp(_f64(rec2)), fmt2, _dp(R), _dp(T))
    return R, T


def pose_decode(rec, fmt=POSE_RT12):
    rec = _f64(rec)
    R = np.empty((3, 3)); T = np.empty(3); q = np.empty(4)
    lib().orc_pose_decode(_dp(rec), fmt, _dp(R), _dp(T), _dp(q))
    return R, T, q
