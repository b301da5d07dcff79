"""ctypes binding of libmptcc.so (include/mptcc.h) -- the only way Python reaches the CUDA path.

There is no fallback: if the shared library is missing, ``load()`` raises, and every ``check_*`` call
raises ``MptccError`` when no sm_100 device is usable.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmptcc.so")
# kernel A/B experiments only (tools/build_variants.sh): another build of the same library
if os.environ.get("MPTCC_LIB"):
    LIB_PATH = os.path.abspath(os.environ["MPTCC_LIB"])

POSE_RT12 = 0
POSE_TQ7 = 1
POSE_T3 = 2       # x y z: Transform3f(Vec3f)
POSE_XYZCS = 3    # x y z cos sin: planar rotation of the blimp / snake
REC_LEN = {0: 12, 1: 7, 2: 3, 3: 5}

OK, ERR_INVALID, ERR_STATE, ERR_NO_DEVICE, ERR_CUDA, ERR_NOMEM, ERR_LIMIT = range(7)
_ERR_NAMES = {1: "INVALID", 2: "STATE", 3: "NO_DEVICE", 4: "CUDA", 5: "NOMEM", 6: "LIMIT"}

# every symbol include/mptcc.h declares (tests check that the library exports exactly these)
EXPORTS = [
    "mptcc_create", "mptcc_destroy", "mptcc_last_error", "mptcc_num_devices", "mptcc_set_environment",
    "mptcc_set_agent", "mptcc_check_poses", "mptcc_submit_poses", "mptcc_wait", "mptcc_check_segments", "mptcc_check_omni_edges", "mptcc_check_poses_device",
    "mptcc_synchronize", "mptcc_alloc_pinned", "mptcc_free_pinned", "mptcc_enable_counters",
    "mptcc_get_counters", "mptcc_reset_counters", "mptcc_last_kernel_ms", "mptcc_last_phase_ms", "mptcc_set_option", "mptcc_get_info",
    "mptcc_bvh_build", "mptcc_bvh_free", "mptcc_bvh_num_nodes", "mptcc_bvh_depth", "mptcc_bvh_get_nodes",
    "mptcc_bvh_get_primitive_indices", "mptcc_bvh_get_local_aabb", "mptcc_bvh_get_quantised_nodes", "mptcc_bvh_get_exact_arrays",
    "mptcc_bvh_get_certificate_data",
    "mptcc_debug_create_host_only", "mptcc_debug_grid_lookup", "mptcc_debug_tree_lookup",
]


class MptccError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__("mptcc error %d (%s)%s" % (code, _ERR_NAMES.get(code, "?"), (": " + msg) if msg else ""))


class Mesh(C.Structure):
    _fields_ = [("xyz", C.POINTER(C.c_double)), ("nv", C.c_uint32), ("tri", C.POINTER(C.c_uint32)), ("nt", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("n_pose_sets", "n_candidates", "n_bv_tests", "n_leaf_tests", "n_exact_tri",
                                          "n_path_checks", "n_kernel_launches")]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


_LIB = None


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError("libmptcc.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C motionplanningtoolkit_b200/csrc`" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, dp, u32p, i32p = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    L.mptcc_create.restype = C.c_int
    L.mptcc_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(C.c_int)]
    L.mptcc_destroy.restype = None
    L.mptcc_destroy.argtypes = [vp]
    L.mptcc_last_error.restype = C.c_char_p
    L.mptcc_last_error.argtypes = [vp]
    L.mptcc_num_devices.restype = C.c_int
    L.mptcc_num_devices.argtypes = [vp]
    L.mptcc_set_environment.restype = C.c_int
    L.mptcc_set_environment.argtypes = [vp, C.c_int, C.POINTER(Mesh), dp]
    L.mptcc_set_agent.restype = C.c_int
    L.mptcc_set_agent.argtypes = [vp, C.c_int, C.POINTER(Mesh)]
    L.mptcc_check_poses.restype = C.c_int
    L.mptcc_check_poses.argtypes = [vp, C.c_uint64, vp, C.c_int, vp]
    L.mptcc_check_segments.restype = C.c_int
    L.mptcc_check_segments.argtypes = [vp, C.c_uint64, C.POINTER(C.c_uint64), vp, C.c_int, C.POINTER(C.c_uint8)]
    L.mptcc_check_omni_edges.restype = C.c_int
    L.mptcc_check_omni_edges.argtypes = [vp, C.c_uint64, vp, C.c_double, vp]
    L.mptcc_submit_poses.restype = C.c_int
    L.mptcc_submit_poses.argtypes = [vp, C.c_uint64, vp, C.c_int, vp, C.POINTER(C.c_uint64)]
    L.mptcc_wait.restype = C.c_int
    L.mptcc_wait.argtypes = [vp, C.c_uint64]
    L.mptcc_check_poses_device.restype = C.c_int
    L.mptcc_check_poses_device.argtypes = [vp, C.c_int, C.c_uint64, vp, C.c_int, vp, vp]
    L.mptcc_synchronize.restype = C.c_int
    L.mptcc_synchronize.argtypes = [vp]
    L.mptcc_alloc_pinned.restype = vp
    L.mptcc_alloc_pinned.argtypes = [C.c_size_t]
    L.mptcc_free_pinned.restype = None
    L.mptcc_free_pinned.argtypes = [vp]
    L.mptcc_enable_counters.restype = C.c_int
    L.mptcc_enable_counters.argtypes = [vp, C.c_int]
    L.mptcc_get_counters.restype = C.c_int
    L.mptcc_get_counters.argtypes = [vp, C.POINTER(Counters)]
    L.mptcc_reset_counters.restype = C.c_int
    L.mptcc_reset_counters.argtypes = [vp]
    L.mptcc_last_kernel_ms.restype = C.c_int
    L.mptcc_last_kernel_ms.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
    L.mptcc_last_phase_ms.restype = C.c_int
    L.mptcc_last_phase_ms.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
    L.mptcc_get_info.restype = C.c_int
    L.mptcc_get_info.argtypes = [vp, C.c_char_p, C.POINTER(C.c_long)]
    L.mptcc_set_option.restype = C.c_int
    L.mptcc_set_option.argtypes = [vp, C.c_char_p, C.c_long]
    L.mptcc_bvh_build.restype = C.c_int
    L.mptcc_bvh_build.argtypes = [C.POINTER(Mesh), C.POINTER(vp)]
    L.mptcc_bvh_free.restype = None
    L.mptcc_bvh_free.argtypes = [vp]
    L.mptcc_bvh_num_nodes.restype = C.c_uint32
    L.mptcc_bvh_num_nodes.argtypes = [vp]
    L.mptcc_bvh_depth.restype = C.c_uint32
    L.mptcc_bvh_depth.argtypes = [vp]
    L.mptcc_bvh_get_nodes.restype = None
    L.mptcc_bvh_get_nodes.argtypes = [vp, dp, dp, dp, i32p, i32p, i32p]
    L.mptcc_bvh_get_primitive_indices.restype = None
    L.mptcc_bvh_get_primitive_indices.argtypes = [vp, u32p]
    L.mptcc_debug_create_host_only.restype = C.c_int
    L.mptcc_debug_create_host_only.argtypes = [C.POINTER(vp)]
    for fn in (L.mptcc_debug_grid_lookup, L.mptcc_debug_tree_lookup):
        fn.restype = C.c_int
        fn.argtypes = [vp, dp, dp, C.c_uint32, u32p]
    L.mptcc_bvh_get_certificate_data.restype = C.c_int
    L.mptcc_bvh_get_certificate_data.argtypes = [vp, C.POINTER(C.c_float), dp]
    L.mptcc_bvh_get_exact_arrays.restype = C.c_int
    L.mptcc_bvh_get_exact_arrays.argtypes = [vp, C.c_int, dp, u32p, u32p]
    L.mptcc_bvh_get_quantised_nodes.restype = C.c_int
    L.mptcc_bvh_get_quantised_nodes.argtypes = [vp, C.c_int, u32p, u32p, dp]
    L.mptcc_bvh_get_local_aabb.restype = None
    L.mptcc_bvh_get_local_aabb.argtypes = [vp, dp, dp, dp]
    _LIB = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class _MeshHolder:
    """Keeps the numpy arrays alive next to the C struct array."""

    def __init__(self, meshes):
        self.arrays = []
        self.c = (Mesh * max(1, len(meshes)))()
        seen = {}
        for i, (xyz, tri) in enumerate(meshes):
            key = (id(xyz), id(tri))
            if key not in seen:       # identical python objects -> identical pointers -> shared device model
                x = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
                t = np.ascontiguousarray(tri, dtype=np.uint32).reshape(-1, 3)
                seen[key] = (x, t)
                self.arrays.append((x, t))
            x, t = seen[key]
            self.c[i].xyz = _dp(x)
            self.c[i].nv = x.shape[0]
            self.c[i].tri = t.ctypes.data_as(C.POINTER(C.c_uint32))
            self.c[i].nt = t.shape[0]


class PinnedArray:
    """numpy view over cudaMallocHost memory (mptcc_alloc_pinned)."""

    def __init__(self, shape, dtype):
        self.lib = load()
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        self.ptr = self.lib.mptcc_alloc_pinned(n)
        if not self.ptr:
            raise MptccError(ERR_CUDA, "cudaMallocHost failed")
        buf = (C.c_char * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.mptcc_free_pinned(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Checker:
    """One mptcc context: environment + agent + batched validity checks on 1..8 B200s."""

    def __init__(self, n_devices=1, device_ids=None):
        self.lib = load()
        h = C.c_void_p()
        ids = None
        if device_ids is not None:
            ids = (C.c_int * len(device_ids))(*device_ids)
            n_devices = len(device_ids)
        rc = self.lib.mptcc_create(C.byref(h), n_devices, ids)
        if rc != OK:
            raise MptccError(rc, "mptcc_create: no usable sm_100 CUDA device" if rc == ERR_NO_DEVICE else "mptcc_create")
        self.h = h
        self.n_links = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.mptcc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise MptccError(rc, self.lib.mptcc_last_error(self.h).decode())

    def set_environment(self, meshes, pose7=(0, 0, 0, 1, 0, 0, 0)):
        """meshes: list of (xyz, tri) -- one collision object each; pose7: one x y z qw qx qy qz for all
        objects (the reference's "Environment Location") or one row per object."""
        mh = _MeshHolder(meshes)
        p = np.ascontiguousarray(pose7, dtype=np.float64)
        if p.ndim == 1:
            p = np.tile(p, (max(1, len(meshes)), 1))
        p = np.ascontiguousarray(p)
        self._check(self.lib.mptcc_set_environment(self.h, len(meshes), mh.c, _dp(p)))

    def set_agent(self, link_meshes):
        mh = _MeshHolder(link_meshes)
        self._check(self.lib.mptcc_set_agent(self.h, len(link_meshes), mh.c))
        self.n_links = len(link_meshes)

    def _rec(self, fmt):
        return REC_LEN[fmt] * self.n_links

    def check_poses(self, poses, fmt=POSE_RT12, out=None):
        """poses: float64 [n, n_links, 12|7] (host).  Returns uint32 verdict words (bit=1: collides)."""
        p = np.ascontiguousarray(poses, dtype=np.float64)
        n = p.size // max(1, self._rec(fmt))
        bits = out if out is not None else np.zeros((n + 31) // 32, np.uint32)
        self._check(self.lib.mptcc_check_poses(self.h, n, p.ctypes.data, fmt, bits.ctypes.data))
        return bits

    def submit_poses(self, poses, out, fmt=POSE_RT12):
        """check_poses without the final wait: returns a ticket for wait().  `poses` (contiguous float64) and
        `out` (uint32 verdict words) must stay alive until then; pinned arrays make the copies asynchronous."""
        if not (isinstance(poses, np.ndarray) and poses.dtype == np.float64 and poses.flags.c_contiguous):
            raise ValueError("submit_poses needs a C-contiguous float64 array (it is read after the call returns)")
        n = poses.size // max(1, self._rec(fmt))
        if out.dtype != np.uint32 or out.size < (n + 31) // 32:
            raise ValueError("out: uint32[ceil(n/32)]")
        t = C.c_uint64(0)
        self._check(self.lib.mptcc_submit_poses(self.h, n, poses.ctypes.data, fmt, out.ctypes.data, C.byref(t)))
        return t.value

    def wait(self, ticket=0):
        """Blocks until `ticket` (0 = everything submitted) and all earlier submits are complete."""
        self._check(self.lib.mptcc_wait(self.h, ticket))

    def check_segments(self, seg_begin, poses, fmt=POSE_RT12):
        sb = np.ascontiguousarray(seg_begin, dtype=np.uint64)
        p = np.ascontiguousarray(poses, dtype=np.float64)
        safe = np.zeros(len(sb) - 1, np.uint8)
        self._check(self.lib.mptcc_check_segments(self.h, len(sb) - 1, sb.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                  p.ctypes.data, fmt, safe.ctypes.data_as(C.POINTER(C.c_uint8))))
        return safe.astype(bool)

    def check_omni_edges(self, start_end, dt):
        """start_end: float64 [n, 6]; poses are interpolated on the device (Omnidirectional::getPoses)."""
        e = np.ascontiguousarray(start_end, dtype=np.float64).reshape(-1, 6)
        safe = np.zeros(len(e), np.uint8)
        self._check(self.lib.mptcc_check_omni_edges(self.h, len(e), e.ctypes.data, float(dt), safe.ctypes.data))
        return safe.astype(bool)

    def check_poses_device(self, n, d_poses_ptr, d_bits_ptr, fmt=POSE_RT12, slot=0, stream=None):
        """Device-resident buffers (raw pointers, e.g. torch.Tensor.data_ptr()); asynchronous."""
        self._check(self.lib.mptcc_check_poses_device(self.h, slot, n, d_poses_ptr, fmt, d_bits_ptr, stream))

    def synchronize(self):
        self._check(self.lib.mptcc_synchronize(self.h))

    def enable_counters(self, on=True):
        self._check(self.lib.mptcc_enable_counters(self.h, int(bool(on))))

    def counters(self):
        c = Counters()
        self._check(self.lib.mptcc_get_counters(self.h, C.byref(c)))
        return c.as_dict()

    def reset_counters(self):
        self._check(self.lib.mptcc_reset_counters(self.h))

    def last_kernel_ms(self, slot=0):
        ms = C.c_float()
        self._check(self.lib.mptcc_last_kernel_ms(self.h, slot, C.byref(ms)))
        return ms.value

    def last_phase_ms(self, slot=0):
        """(broadphase, narrowphase, exact pass, fall-back) ms of the last split-pipeline launch (option time_phases)."""
        ms = (C.c_float * 4)()
        self._check(self.lib.mptcc_last_phase_ms(self.h, slot, ms))
        return tuple(ms)

    def info(self, name):
        v = C.c_long(0)
        self._check(self.lib.mptcc_get_info(self.h, name.encode(), C.byref(v)))
        return int(v.value)

    def set_option(self, name, value):
        self._check(self.lib.mptcc_set_option(self.h, name.encode(), int(value)))


class HostScene(Checker):
    """A context WITHOUT devices (mptcc_debug_create_host_only): set_environment / set_agent build the host side of the
    scene; grid_lookup / tree_lookup return what the broadphase structures answer for one pose box.  Tests only."""

    def __init__(self):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.mptcc_debug_create_host_only(C.byref(h))
        if rc != OK:
            raise MptccError(rc, "mptcc_debug_create_host_only")
        self.h = h
        self.n_links = 0

    def _lookup(self, fn, box6, cap=4096):
        box = np.ascontiguousarray(box6, dtype=np.float64)
        out = np.zeros((cap, 6))
        listed = C.c_uint32(0)
        n = fn(self.h, _dp(box), _dp(out), cap, C.byref(listed))
        return n, out[:max(0, min(n, cap))], listed.value

    def grid_lookup(self, box6):
        """(count or -1 / -2, world boxes of the overlapping listed objects [count, 6], objects listed for the cell)"""
        return self._lookup(self.lib.mptcc_debug_grid_lookup, box6)

    def tree_lookup(self, box6):
        return self._lookup(self.lib.mptcc_debug_tree_lookup, box6)


class HostBvh:
    """Host-side OBB tree of libmptcc (no GPU needed)."""

    def __init__(self, xyz, tri):
        self.lib = load()
        self._mh = _MeshHolder([(xyz, tri)])
        h = C.c_void_p()
        rc = self.lib.mptcc_bvh_build(self._mh.c, C.byref(h))
        if rc != OK:
            raise MptccError(rc, "mptcc_bvh_build")
        self.h = h
        self.nt = self._mh.c[0].nt

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.mptcc_bvh_free(self.h)
            self.h = None

    @property
    def num_nodes(self):
        return self.lib.mptcc_bvh_num_nodes(self.h)

    @property
    def depth(self):
        return self.lib.mptcc_bvh_depth(self.h)

    def nodes(self):
        n = self.num_nodes
        axis = np.empty((n, 3, 3)); To = np.empty((n, 3)); ext = np.empty((n, 3))
        fc = np.empty(n, np.int32); fp = np.empty(n, np.int32); npr = np.empty(n, np.int32)
        i32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        self.lib.mptcc_bvh_get_nodes(self.h, _dp(axis), _dp(To), _dp(ext), i32(fc), i32(fp), i32(npr))
        return dict(axis=axis, To=To, extent=ext, first_child=fc, first_primitive=fp, num_primitives=npr)

    def quantised_nodes(self, breadth_first=False):
        """(words [n, 8] uint32, device_id [n], ctr [3], ok): the 32-byte traversal nodes the device would get."""
        n = self.num_nodes
        words = np.zeros((n, 8), np.uint32); dev = np.zeros(n, np.uint32); ctr = np.zeros(3)
        u32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
        rc = self.lib.mptcc_bvh_get_quantised_nodes(self.h, int(breadth_first), u32(words), u32(dev), _dp(ctr))
        if rc < 0:
            raise MptccError(rc, "mptcc_bvh_get_quantised_nodes")
        return words, dev, ctr, bool(rc)

    def exact_arrays(self, breadth_first=False):
        """(node64 [n, 16], anc [num_tris, depth]): what the exact pass reads for this model."""
        u32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
        depth = C.c_uint32(0)
        self.lib.mptcc_bvh_get_exact_arrays(self.h, int(breadth_first), None, None, C.byref(depth))
        node64 = np.zeros((self.num_nodes, 16)); anc = np.zeros((self.nt, depth.value), np.uint32)
        if self.lib.mptcc_bvh_get_exact_arrays(self.h, int(breadth_first), _dp(node64), u32(anc), C.byref(depth)) != 0:
            raise MptccError(ERR_INVALID, "mptcc_bvh_get_exact_arrays")
        return node64, anc

    def certificate_data(self):
        """(mu_chain [num_tris], dict(ctr, cmax1, esum, r2)): what the path certificate uses of this model."""
        mu = np.zeros(self.nt, np.float32); m6 = np.zeros(6)
        if self.lib.mptcc_bvh_get_certificate_data(self.h, mu.ctypes.data_as(C.POINTER(C.c_float)), _dp(m6)) != 0:
            raise MptccError(ERR_INVALID, "mptcc_bvh_get_certificate_data")
        return mu, dict(ctr=m6[:3].copy(), cmax1=m6[3], esum=m6[4], r2=m6[5])

    def primitive_indices(self):
        out = np.empty(self.nt, np.uint32)
        self.lib.mptcc_bvh_get_primitive_indices(self.h, out.ctypes.data_as(C.POINTER(C.c_uint32)))
        return out

    def local_aabb(self):
        box = np.empty(6); ctr = np.empty(3); r = C.c_double()
        self.lib.mptcc_bvh_get_local_aabb(self.h, _dp(box), _dp(ctr), C.byref(r))
        return box, ctr, r.value
