#!/usr/bin/env python3
"""Generate tests/golden/*.npz: seeded pose batches with the oracle's verdicts and BVH digests.

The reference ships no golden vectors and FCL is not installable here (parity unpinned), so these
fixtures pin the ORACLE ITSELF (against accidental changes) and give the GPU tests a box-independent
expectation.  The analytic known answers in tests/test_oracle_kat.py are the independent check.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from motionplanningtoolkit_b200 import meshio, scenes  # noqa: E402
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def digest(model):
    nd = model.nodes()
    h = hashlib.sha256()
    for k in ("axis", "To", "extent", "first_child", "first_primitive", "num_primitives"):
        h.update(np.ascontiguousarray(nd[k]).tobytes())
    h.update(model.primitive_indices().tobytes())
    return h.hexdigest()


ONLY = sys.argv[1:]  # optional: names of the cases to (re)write


def case(name, env, links, poses, fmt=O.POSE_RT12, pose7=(0, 0, 0, 1, 0, 0, 0), self_collision=False):
    if ONLY and name not in ONLY:
        return
    models = {}
    om = [models.setdefault(id(m), O.Model(*m)) for m in env]
    ol = [models.setdefault(id(m), O.Model(*m)) for m in links]
    oenv = O.Environment(om, pose7)
    ctr = O.Counters()
    bits = (O.verdicts_self if self_collision else O.verdicts)(oenv, ol, poses, fmt, 8, ctr)
    n = poses.shape[0]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), poses=poses, bits=bits, fmt=fmt, pose7=np.asarray(pose7, float),
                        self_collision=int(self_collision), pinned_by=np.array(""),  # set by tools/pin_golden_with_fcl.py

                        n_bv_tests=ctr.n_bv_tests, n_leaf_tests=ctr.n_leaf_tests,
                        env_digest=np.array([digest(m) for m in om[:8]]), link_digest=np.array([digest(m) for m in ol[:1]]))
    print("%-28s n=%6d collide=%.3f bv=%d leaf=%d" % (name, n, O.unpack_bits(bits, n).mean(), ctr.n_bv_tests, ctr.n_leaf_tests))


def main():
    os.makedirs(OUT, exist_ok=True)
    box = meshio.load_mptm("unit_box")[0]
    blimp = meshio.load_mptm("blimp")
    divider = meshio.environment_submeshes(meshio.load_mptm("divider"))
    model = meshio.environment_submeshes(meshio.load_mptm("model"))
    rng = scenes.SplitMix64(0x601D)
    # C1: omnidirectional.inst geometry, translation-only poses incl. exact-touch cases
    e = 2.0 ** -52
    kat = np.array([[0, 0, 0], [1, 0, 0], [1 + e, 0, 0], [1 - e, 0, 0], [0, -1, 0], [0, 0, 1 + e], [1, 1, 1], [1, 1, 1 + e]], float)
    T = np.concatenate([kat, np.stack([rng.uniform(3000, -1.6, 1.6) for _ in range(3)], 1)])
    case("c1_box_translations", [box], [box], scenes.translation_poses(T))
    case("c1_box_rotations", [box], [box], scenes.random_poses_rt12(rng, 3000, (-1.8,) * 3, (1.8,) * 3))
    case("c1_box_quaternions", [box], [box], scenes.random_poses_tq7(rng, 2000, (-1.8,) * 3, (1.8,) * 3), O.POSE_TQ7)
    # C3: blimp (last sub-mesh, as the reference loads it; and merged) vs divider / model
    last, merged = meshio.agent_submesh(blimp, "last"), meshio.agent_submesh(blimp, "merge")
    case("c3_blimp_last_divider", divider, [last], scenes.random_poses_near_objects(rng, 4000, divider, 1.0))
    case("c3_blimp_merge_divider", divider, [merged], scenes.random_poses_near_objects(rng, 3000, divider, 2.0))
    case("c3_blimp_merge_model", model, [merged], scenes.random_poses_near_objects(rng, 3000, model, 2.0))
    q = np.array([0.8, 0.2, -0.4, 0.4]); q /= np.linalg.norm(q)
    Renv, Tenv = scenes.quat_to_matrix(q[None])[0], np.array([10., -20., 5.])
    loc = scenes.random_poses_near_objects(rng, 3000, model, 2.0)      # poses in the environment's own frame ...
    Rl, Tl = loc[:, :9].reshape(-1, 3, 3), loc[:, 9:]
    world = scenes.poses_rt12(Renv @ Rl, Tl @ Renv.T + Tenv)           # ... moved by the Environment Location
    case("c3_blimp_last_model_rotated_env", model, [last], world, pose7=np.concatenate([Tenv, q]))
    # C4: snake, 11 links of the same box
    poses = np.stack([scenes.random_poses_rt12(rng, 1500, (-4, -4, -1), (4, 4, 1)) for _ in range(11)], axis=1)
    case("c4_snake_11_links", [box], [box] * 11, poses)
    # N4: checkSelfCollision = true, four independent box links (self-collision OR environment)
    rng4 = scenes.SplitMix64(0x601E)
    poses = np.stack([scenes.random_poses_rt12(rng4, 3000, (-2.2,) * 3, (2.2,) * 3) for _ in range(4)], axis=1)
    case("n4_self_collision_4_links", [box], [box] * 4, poses, self_collision=True)


if __name__ == "__main__":
    main()
