"""Shared scene builders for the tests: the same meshes go to the CPU oracle and to libmptcc."""
import numpy as np

from motionplanningtoolkit_b200 import capi, meshio, scenes
from oracle import oracle as O

IDENT7 = (0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0)


class Scene:
    def __init__(self, env_meshes, link_meshes, env_pose7=IDENT7):
        self.env_meshes = list(env_meshes)
        self.link_meshes = list(link_meshes)
        self.env_pose7 = np.asarray(env_pose7, dtype=np.float64)
        self._omodels = {}
        self.oenv = O.Environment([self._omodel(m) for m in self.env_meshes], self.env_pose7)
        self.olinks = [self._omodel(m) for m in self.link_meshes]

    def _omodel(self, mesh):
        k = id(mesh)
        if k not in self._omodels:
            self._omodels[k] = O.Model(*mesh)
        return self._omodels[k]

    def checker(self, n_devices=1, **options):
        c = capi.Checker(n_devices)
        for k, v in options.items():
            c.set_option(k, v)
        c.set_environment(self.env_meshes, self.env_pose7)
        c.set_agent(self.link_meshes)
        return c

    def oracle_bits(self, poses, fmt=O.POSE_RT12, n_threads=8, counters=None):
        return O.verdicts(self.oenv, self.olinks, poses, fmt, n_threads, counters)

    def oracle_bits_self(self, poses, fmt=O.POSE_RT12, n_threads=8, counters=None):
        """checkSelfCollision = true: self-collision of the pose-set OR environment collision."""
        return O.verdicts_self(self.oenv, self.olinks, poses, fmt, n_threads, counters)

    def brute_bits(self, poses, fmt=O.POSE_RT12):
        return O.verdicts_bruteforce(self.oenv, self.olinks, poses, fmt)

    def brute_bits_fast(self, poses, fmt=O.POSE_RT12, n_threads=8):
        return O.verdicts_bruteforce_fast(self.oenv, self.olinks, poses, fmt, n_threads)


def unit_box():
    return meshio.load_mptm("unit_box")[0]


def env_bbox(env_meshes, pad=0.0):
    allv = np.concatenate([m[0] for m in env_meshes])
    return allv.min(0) - pad, allv.max(0) + pad


def assert_bits_equal(got, want, n, what=""):
    g, w = O.unpack_bits(got, n), O.unpack_bits(want, n)
    bad = np.nonzero(g != w)[0]
    assert bad.size == 0, "%s: %d/%d verdicts differ, first at %s (gpu=%s oracle=%s)" % (
        what, bad.size, n, bad[:8], g[bad[:8]], w[bad[:8]])
