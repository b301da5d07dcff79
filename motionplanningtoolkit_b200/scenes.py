"""Synthetic inputs for tests and bench.py: seeded pose batches and procedural environments.

Nothing here is on the product path; these are the "synthetic pose batches" BASELINE.json asks for.
RNG is splitmix64 (seeds fixed in BASELINE.md section 4) so that every box generates identical inputs.
"""
import numpy as np

from . import meshio

SEED_C1, SEED_C2, SEED_C3, SEED_C4, SEED_C5, SEED_ENV = (0x5EED0001, 0x5EED0002, 0x5EED0003, 0x5EED0004,
                                                         0x5EED0005, 0x5EEDE1F0)


class SplitMix64:
    """Vectorised splitmix64: element i of a draw of n is the (counter+i)-th output of the stream."""

    def __init__(self, seed):
        self.state = np.uint64(seed)

    def next_u64(self, n):
        n = int(n)
        with np.errstate(over="ignore"):
            idx = np.arange(1, n + 1, dtype=np.uint64)
            z = self.state + idx * np.uint64(0x9E3779B97F4A7C15)
            self.state = self.state + np.uint64(n) * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))

    def uniform(self, n, lo=0.0, hi=1.0):
        u = (self.next_u64(n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        return lo + (hi - lo) * u

    def integers(self, n, hi):
        return (self.next_u64(n) % np.uint64(hi)).astype(np.int64)


# ---------------------------------------------------------------------------------------------------
# poses
# ---------------------------------------------------------------------------------------------------

def random_quaternions(rng, n):
    """PRMLite::getRandomQuaternion (/root/reference/discretizations/workspace/prmlite.hpp:246-255):
    fcl::Quaternion3f(a,b,c,d) is (w,x,y,z)."""
    u1, u2, u3 = rng.uniform(n), rng.uniform(n), rng.uniform(n)
    return np.stack([np.sqrt(1 - u1) * np.sin(2 * np.pi * u2), np.sqrt(1 - u1) * np.cos(2 * np.pi * u2),
                     np.sqrt(u1) * np.sin(2 * np.pi * u3), np.sqrt(u1) * np.cos(2 * np.pi * u3)], axis=1)


def quat_to_matrix(q):
    """Rotation matrices of unit quaternions (w,x,y,z) -- only used to PRODUCE RT12 inputs."""
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((len(q), 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - w * z); R[:, 0, 2] = 2 * (x * z + w * y)
    R[:, 1, 0] = 2 * (x * y + w * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - w * x)
    R[:, 2, 0] = 2 * (x * z - w * y); R[:, 2, 1] = 2 * (y * z + w * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def poses_rt12(R, T):
    """[n,12] records from rotation matrices [n,3,3] and translations [n,3]."""
    n = len(T)
    out = np.empty((n, 12))
    out[:, :9] = np.asarray(R).reshape(n, 9)
    out[:, 9:] = T
    return out


def translation_poses(T):
    """fcl::Transform3f(Vec3f): identity rotation (Omnidirectional::getPoses)."""
    T = np.asarray(T, dtype=np.float64).reshape(-1, 3)
    return poses_rt12(np.tile(np.eye(3), (len(T), 1, 1)), T)


def random_poses_rt12(rng, n, lo, hi):
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    T = np.stack([rng.uniform(n, lo[k], hi[k]) for k in range(3)], axis=1)
    return poses_rt12(quat_to_matrix(random_quaternions(rng, n)), T)


def random_poses_tq7(rng, n, lo, hi):
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    T = np.stack([rng.uniform(n, lo[k], hi[k]) for k in range(3)], axis=1)
    return np.concatenate([T, random_quaternions(rng, n)], axis=1)


def random_poses_near_objects(rng, n, env_meshes, pad, fmt="rt12"):
    """Translations drawn inside the (padded) AABB of a randomly chosen environment object, uniform
    rotations: exercises the narrowphase instead of empty space."""
    lo = np.array([m[0].min(0) for m in env_meshes]) - pad
    hi = np.array([m[0].max(0) for m in env_meshes]) + pad
    k = rng.integers(n, len(env_meshes))
    T = np.stack([lo[k, a] + (hi[k, a] - lo[k, a]) * rng.uniform(n) for a in range(3)], axis=1)
    q = random_quaternions(rng, n)
    if fmt == "tq7":
        return np.concatenate([T, q], axis=1)
    return poses_rt12(quat_to_matrix(q), T)


def omnidirectional_edge_poses(start, end, dt):
    """Omnidirectional::getPoses (/root/reference/agents/omnidirectional.hpp:202-247): translation-only
    interpolation; returns [P,3] translations."""
    start, end = np.asarray(start, float), np.asarray(end, float)
    d = end - start
    dist = float(np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
    iterations = int(dist / dt) if dist > 0 else 0
    if iterations < 1:
        return np.stack([start, end])
    step = dt / dist
    pts = [start + (step * float(i)) * d for i in range(iterations)]
    if float(iterations) * dt < dist:
        pts.append(end)
    return np.stack(pts)


def omnidirectional_edges_poses(starts, ends, dt):
    """omnidirectional_edge_poses for many edges at once (same arithmetic, vectorised): returns
    (translations [P,3], seg_begin [n+1])."""
    starts, ends = np.asarray(starts, float).reshape(-1, 3), np.asarray(ends, float).reshape(-1, 3)
    d = ends - starts
    dist = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = np.where(dist > 0, dist / dt, 0.0)
        it = np.where(ratio >= 1.0, np.minimum(ratio, 4294967295.0), 0.0).astype(np.int64)
        extra = (it.astype(np.float64) * dt < dist) & (it >= 1)
        cnt = np.where(it < 1, 2, it + extra)
        seg = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint64)
        e = np.repeat(np.arange(len(cnt)), cnt)
        i = np.arange(int(seg[-1])) - np.repeat(seg[:-1].astype(np.int64), cnt)
        step = np.where(dist > 0, dt / dist, 0.0)
        s = (step[e] * i.astype(np.float64))[:, None]
        pts = starts[e] + s * d[e]
    short = it[e] < 1
    pts[short & (i == 0)] = starts[e][short & (i == 0)]
    pts[short & (i == 1)] = ends[e][short & (i == 1)]
    last = (~short) & (i >= it[e])
    pts[last] = ends[e][last]
    return pts, seg


# ---------------------------------------------------------------------------------------------------
# procedural meshes
# ---------------------------------------------------------------------------------------------------

def icosphere(subdiv=2, radius=1.0):
    t = (1.0 + 5 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, float) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache, nf = {}, []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[k] = len(v) - 1
            return cache[k]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.array(v) * radius, np.array(f, dtype=np.uint32)


def grid_box(n=4, size=(1.0, 1.0, 1.0)):
    """Axis-aligned box whose faces are n x n quads (2 triangles each): 12 n^2 triangles."""
    verts, tris = [], []
    sx, sy, sz = [s / 2.0 for s in size]
    lin = np.linspace(-1, 1, n + 1)
    for axis in range(3):
        for sign in (-1, 1):
            base = len(verts)
            for i in range(n + 1):
                for j in range(n + 1):
                    p = [0, 0, 0]
                    p[axis] = sign
                    p[(axis + 1) % 3] = lin[i]
                    p[(axis + 2) % 3] = lin[j]
                    verts.append((p[0] * sx, p[1] * sy, p[2] * sz))
            for i in range(n):
                for j in range(n):
                    a = base + i * (n + 1) + j
                    b, c, d = a + 1, a + (n + 1), a + (n + 1) + 1
                    tris += [(a, b, d), (a, d, c)] if sign > 0 else [(a, d, b), (a, c, d)]
    return np.array(verts, float), np.array(tris, dtype=np.uint32)


def transform_mesh(xyz, R, T):
    return xyz @ np.asarray(R).T + np.asarray(T)


def synthetic_environment(n_objects=2000, extent=200.0, seed=SEED_ENV):
    """n_objects randomly placed / oriented / scaled obstacles, vertices baked into world
    coordinates (every object's pose is the identity, as in the reference where all sub-meshes share one
    "Environment Location").  Mix: 50 % gridded boxes (432 triangles), 30 % icospheres (320), 20 %
    icospheres (1280): 568 triangles per object on average, ~1.14 M triangles for 2000 objects.
    Returns a list of (xyz, tri)."""
    rng = SplitMix64(seed)
    shapes = [grid_box(6), icosphere(2), icosphere(3)]
    shapes[0] = (shapes[0][0] * 2.0, shapes[0][1])
    objs = []
    pick = rng.uniform(n_objects)
    kinds = np.where(pick < 0.5, 0, np.where(pick < 0.8, 1, 2))
    R = quat_to_matrix(random_quaternions(rng, n_objects))
    T = np.stack([rng.uniform(n_objects, -extent / 2, extent / 2) for _ in range(3)], axis=1)
    scale = np.stack([rng.uniform(n_objects, 1.0, 5.0) for _ in range(3)], axis=1)
    for i in range(n_objects):
        v, t = shapes[int(kinds[i])]
        objs.append((np.ascontiguousarray(transform_mesh(v * scale[i], R[i], T[i])), t))
    return objs


def terrain(nx=708, ny=708, extent=200.0, height=6.0, seed=SEED_ENV):
    """Single-object heightfield with 2*nx*ny triangles (default ~1.0 M)."""
    rng = SplitMix64(seed + 1)
    xs = np.linspace(-extent / 2, extent / 2, nx + 1)
    ys = np.linspace(-extent / 2, extent / 2, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    Z = np.zeros_like(X)
    for _ in range(24):
        fx, fy = rng.uniform(1, 0.02, 0.4)[0], rng.uniform(1, 0.02, 0.4)[0]
        ph, am = rng.uniform(1, 0, 6.283)[0], rng.uniform(1, 0.1, 1.0)[0]
        Z += am * np.sin(fx * X + fy * Y + ph)
    Z *= height / max(1e-9, np.abs(Z).max())
    v = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    idx = np.arange((nx + 1) * (ny + 1)).reshape(nx + 1, ny + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel()
    t = np.concatenate([np.stack([a, b, d], 1), np.stack([a, d, c], 1)]).astype(np.uint32)
    return np.ascontiguousarray(v), t


def fixture(name):
    return meshio.load_mptm(name)


# ---------------------------------------------------------------------------------------------------
# pose producers of the reference's agents (inputs of the hot path; restated verbatim incl. quirks)
# ---------------------------------------------------------------------------------------------------

def _normalize_theta(t):
    """Blimp/SnakeTrailers::normalizeTheta (/root/reference/agents/blimp.hpp:358-360)."""
    return t - 2 * np.pi * np.floor((t + np.pi) / (2 * np.pi))


def _rz_convention(theta, T):
    """R(0,0)=c, R(1,0)=-s, R(0,1)=s, R(1,1)=c (blimp.hpp:350-353, snake_trailers.hpp:427-430)."""
    n = len(theta)
    R = np.tile(np.eye(3), (n, 1, 1))
    c, s = np.cos(theta), np.sin(theta)
    R[:, 0, 0] = c
    R[:, 1, 0] = -s
    R[:, 0, 1] = s
    R[:, 1, 1] = c
    return poses_rt12(R, T)


BLIMP_INST = dict(length=10.0, vmin=-1.0, vmax=5.0, psi_min=-0.785398, psi_max=0.785398, vz_min=-5.0, vz_max=5.0)


def blimp_do_step(s, a, w, z, dt, P=BLIMP_INST):
    """Blimp::doStep (/root/reference/agents/blimp.hpp:293-317); state columns x y z theta v psi vz.
    (theta advances by v*tan(psi)/length WITHOUT dt, as in the reference.)"""
    x, y, zz, th, v, psi, vz = s.T
    n = np.empty_like(s)
    n[:, 0] = x + np.cos(th) * v * dt
    n[:, 1] = y + np.sin(th) * v * dt
    n[:, 3] = _normalize_theta(th + v * np.tan(psi) / P["length"])
    n[:, 2] = zz + vz * dt
    n[:, 4] = np.clip(v + a * dt, P["vmin"], P["vmax"])
    n[:, 5] = np.clip(psi + w * dt, P["psi_min"], P["psi_max"])
    n[:, 6] = np.clip(vz + z * dt, P["vz_min"], P["vz_max"])
    return n


def blimp_edge_poses(rng, n_edges, lo, hi, steps=10, dt=0.1):
    """uniform start states inside [lo,hi], uniform controls (a in [-1,1], w in
    [-0.1745,0.1745], z in [-1,1], blimp.hpp:123), `steps` poses per edge via stateToFCLTransform
    (blimp.hpp:339-356; Blimp::getPoses itself is a stub).  Returns (poses [n*steps,12],
    seg_begin [n+1])."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    s = np.zeros((n_edges, 7))
    for k in range(3):
        s[:, k] = rng.uniform(n_edges, lo[k], hi[k])
    s[:, 3] = rng.uniform(n_edges, -np.pi, np.pi)
    s[:, 4] = rng.uniform(n_edges, BLIMP_INST["vmin"], BLIMP_INST["vmax"])
    s[:, 5] = rng.uniform(n_edges, BLIMP_INST["psi_min"], BLIMP_INST["psi_max"])
    s[:, 6] = rng.uniform(n_edges, BLIMP_INST["vz_min"], BLIMP_INST["vz_max"])
    a, w, z = rng.uniform(n_edges, -1, 1), rng.uniform(n_edges, -0.1745, 0.1745), rng.uniform(n_edges, -1, 1)
    out = np.empty((n_edges, steps, 12))
    for i in range(steps):
        out[:, i] = _rz_convention(s[:, 3], s[:, :3])
        s = blimp_do_step(s, a, w, z, dt)
    return out.reshape(-1, 12), np.arange(n_edges + 1, dtype=np.uint64) * steps


SNAKE_INST = dict(trailers=10, length=1.0, hitch=0.25, vmin=-1.0, vmax=5.0, psi_min=-0.785398, psi_max=0.785398)


def snake_state_poses(states, P=SNAKE_INST):
    """SnakeTrailers::stateToFCLTransforms (/root/reference/agents/snake_trailers.hpp:411-459), verbatim
    including Q4: every trailer sits at (-(L+h), y0, 0) and carries the RELATIVE hitch rotation.
    states: [n, 5 + trailers] = x y theta v psi theta_1..theta_k.  Returns [n, trailers+1, 12]."""
    n, k = len(states), P["trailers"]
    out = np.empty((n, k + 1, 12))
    T = np.zeros((n, 3))
    T[:, 0], T[:, 1] = states[:, 0], states[:, 1]
    out[:, 0] = _rz_convention(states[:, 2], T)
    T = T.copy()
    T[:, 0] = -(P["length"] + P["hitch"])
    th = np.concatenate([states[:, 2:3], states[:, 5:5 + k]], axis=1)   # THETA, THETA+1, ...
    for i in range(1, k + 1):
        out[:, i] = _rz_convention(th[:, i] - th[:, i - 1], T)
    return out


def snake_random_states(rng, n, lo, hi, P=SNAKE_INST):
    k = P["trailers"]
    s = np.zeros((n, 5 + k))
    s[:, 0], s[:, 1] = rng.uniform(n, lo[0], hi[0]), rng.uniform(n, lo[1], hi[1])
    s[:, 2] = rng.uniform(n, -np.pi, np.pi)
    s[:, 3] = rng.uniform(n, P["vmin"], P["vmax"])
    s[:, 4] = rng.uniform(n, P["psi_min"], P["psi_max"])
    for i in range(k):
        s[:, 5 + i] = _normalize_theta(s[:, 2] + rng.uniform(n, -0.6, 0.6))
    return s


def prm_slice_edges(rng, n_milestones, lo, hi, k=10, dt=0.1, existing=None):
    """BASELINE.md C2 / PRM::addMilestone (/root/reference/planners/prm/prm.hpp:334-387): every new
    milestone is connected by a full straight edge (Omnidirectional::steer with fraction clamped to 1) to
    its k nearest EARLIER milestones; each edge is interpolated at dt.  Returns (poses, seg_begin, points)."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    pts = np.stack([rng.uniform(n_milestones, lo[a], hi[a]) for a in range(3)], axis=1)
    allp = pts if existing is None else np.concatenate([existing, pts])
    first = 0 if existing is None else len(existing)
    segs, chunks = [0], []
    for i in range(first, len(allp)):
        if i == 0:
            continue
        d = np.linalg.norm(allp[:i] - allp[i], axis=1)
        for j in np.argsort(d, kind="stable")[:k]:
            e = omnidirectional_edge_poses(allp[i], allp[j], dt)
            chunks.append(e)
            segs.append(segs[-1] + len(e))
    T = np.concatenate(chunks) if chunks else np.zeros((0, 3))
    return translation_poses(T), np.asarray(segs, dtype=np.uint64), allp
