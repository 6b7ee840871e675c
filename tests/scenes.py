"""Deterministic synthetic inputs shared by tests, bench.py and smoke().

Fixture geometry follows the reference's factories (src/mesh.cpp:25-137); scene generators follow
BASELINE.md ("Inputs": splitmix64, seed 0x5EED0000 + config_id, Shoemake quaternions).
Shape specs: ("aabb", min3, max3) | ("obb", half3) | ("sphere", r) | ("hull", verts[n,3]).
"""
from __future__ import annotations

import math

import numpy as np

IDENT = (0.0, 0.0, 0.0, 1.0)  # quaternion memory order x,y,z,w (lin_alg.h:388)


# ---------------------------------------------------------------- fixture meshes (mesh.cpp)
def box_vertices(half):
    hx, hy, hz = (float(h) for h in half)
    return np.array(
        [
            [-hx, -hy, -hz],
            [+hx, -hy, -hz],
            [+hx, +hy, -hz],
            [-hx, +hy, -hz],
            [-hx, -hy, +hz],
            [+hx, -hy, +hz],
            [+hx, +hy, +hz],
            [-hx, +hy, +hz],
        ],
        dtype=np.float64,
    )  # src/mesh.cpp:31-40


def sphere_vertices(radius, stacks=16, sectors=32):
    v = [[0.0, radius, 0.0]]
    for i in range(1, stacks):
        phi = math.pi * float(i) / float(stacks)
        for j in range(sectors):
            theta = 2.0 * math.pi * float(j) / float(sectors)
            v.append(
                [
                    radius * math.sin(phi) * math.cos(theta),
                    radius * math.cos(phi),
                    radius * math.sin(phi) * math.sin(theta),
                ]
            )
    v.append([0.0, -radius, 0.0])
    return np.array(v, dtype=np.float64)  # src/mesh.cpp:71-88 → 482 verts by default


def pyramid_vertices(base_half, height):
    b = float(base_half)
    return np.array(
        [[-b, 0.0, -b], [+b, 0.0, -b], [+b, 0.0, +b], [-b, 0.0, +b], [0.0, float(height), 0.0]],
        dtype=np.float64,
    )  # src/mesh.cpp:119-125


def angle_axis(angle, axis):
    """Eigen AngleAxis → Quaternion (lin_alg.h:569-576): w = cos(a/2), vec = sin(a/2)·axis."""
    ha = 0.5 * float(angle)
    s = math.sin(ha)
    return (s * float(axis[0]), s * float(axis[1]), s * float(axis[2]), math.cos(ha))


def normalized(v):
    v = np.asarray(v, dtype=np.float64)
    return v / math.sqrt(float(v @ v))


# ---------------------------------------------------------------- PRNG
_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


class SplitMix64:
    """Vectorised splitmix64: stream element i is mix(seed + (i+1)·γ)."""

    GAMMA = 0x9E3779B97F4A7C15

    def __init__(self, seed):
        self.state = int(seed) & 0xFFFFFFFFFFFFFFFF

    def next_u64(self, n):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        with np.errstate(over="ignore"):
            z = np.uint64(self.state) + idx * np.uint64(self.GAMMA)
            self.state = int(z[-1]) if n else self.state
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        return z

    def u01(self, *shape):
        n = int(np.prod(shape)) if shape else 1
        x = (self.next_u64(n) >> np.uint64(11)).astype(np.float64) * (2.0**-53)
        return x.reshape(shape) if shape else float(x[0])

    def uniform(self, lo, hi, *shape):
        return lo + (hi - lo) * self.u01(*shape)

    def quats(self, n):
        """Shoemake uniform unit quaternions, x,y,z,w."""
        u = self.u01(n, 3)
        a = np.sqrt(1.0 - u[:, 0])
        b = np.sqrt(u[:, 0])
        t1 = 2.0 * math.pi * u[:, 1]
        t2 = 2.0 * math.pi * u[:, 2]
        return np.stack([a * np.sin(t1), a * np.cos(t1), b * np.sin(t2), b * np.cos(t2)], axis=1)

    def randint(self, n, hi):
        return (self.next_u64(n) % np.uint64(hi)).astype(np.int64)


class Scene:
    """Body arrays + shape specs in the layout pk_bodies_upload / oracle.World.step expect."""

    def __init__(self, shapes, pos, quat, shape_id, flags=None, disp=None):
        self.shapes = list(shapes)
        self.pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        self.quat = np.ascontiguousarray(quat, dtype=np.float64).reshape(-1, 4)
        self.shape_id = np.ascontiguousarray(shape_id, dtype=np.uint32)
        n = len(self.pos)
        self.flags = (
            np.full(n, 2, dtype=np.uint8) if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        )
        self.disp = (
            np.zeros((n, 3)) if disp is None else np.ascontiguousarray(disp, dtype=np.float64).reshape(-1, 3)
        )

    @property
    def n(self):
        return len(self.pos)


# ---------------------------------------------------------------- BASELINE configs (scalable)
def scene_c2(n=100_000, seed=0x5EED0002, extent=None):
    """C2: n randomly posed OBBs; centres U(-L,L)^3, half U(.25,1)^3.  L scales with n so the
    density (≈2 candidate pairs per body at n=1e5, L=50) stays that of the 100 m cube."""
    rng = SplitMix64(seed)
    L = 50.0 * (n / 100_000.0) ** (1.0 / 3.0) if extent is None else extent
    pos = rng.uniform(-L, L, n, 3)
    half = rng.uniform(0.25, 1.0, n, 3)
    quat = rng.quats(n)
    shapes = [("obb", half[i]) for i in range(n)]
    return Scene(shapes, pos, quat, np.arange(n))


def scene_c3(side=100, seed=0x5EED0003, spacing=0.8, jitter=0.2, n_shapes=4096):
    """C3: side³ bodies on a jittered lattice; even ids analytic spheres r~U(.2,.5), odd ids
    OBBs half~U(.2,.5)³.  Shapes are drawn from a table of n_shapes entries (even = sphere,
    odd = box) so the shape table stays small at 1 M bodies."""
    rng = SplitMix64(seed)
    n = side**3
    ii = np.arange(n)
    gx, gy, gz = ii % side, (ii // side) % side, ii // (side * side)
    pos = np.stack([gx, gy, gz], axis=1).astype(np.float64) * spacing
    pos += rng.uniform(-jitter, jitter, n, 3)
    quat = rng.quats(n)
    r = rng.uniform(0.2, 0.5, n_shapes)
    half = rng.uniform(0.2, 0.5, n_shapes, 3)
    shapes = [("sphere", r[k]) if k % 2 == 0 else ("obb", half[k]) for k in range(n_shapes)]
    pick = rng.randint(n, n_shapes // 2)
    shape_id = (2 * pick + (ii % 2)).astype(np.uint32)
    return Scene(shapes, pos, quat, shape_id)


def hull_library(n_hulls=1024, seed=0x5EED0004, sizes=(32, 64, 128, 256)):
    """C4 hull library: points on the unit sphere scaled per axis by U(.5,1.5)·U(.3,.6)."""
    rng = SplitMix64(seed)
    shapes = []
    radii = []
    for k in range(n_hulls):
        V = sizes[k % len(sizes)]
        u = rng.u01(V, 2)
        z = 2.0 * u[:, 0] - 1.0
        t = 2.0 * math.pi * u[:, 1]
        s = np.sqrt(np.maximum(0.0, 1.0 - z * z))
        pts = np.stack([s * np.cos(t), s * np.sin(t), z], axis=1)
        sc = rng.uniform(0.5, 1.5, 3) * rng.uniform(0.3, 0.6)
        pts = pts * sc
        shapes.append(("hull", pts))
        radii.append(float(np.sqrt((pts * pts).sum(axis=1)).max()))
    return shapes, np.array(radii)


def scene_c4(n_pairs=10_000, n_hulls=64, seed=0x5EED0004, sizes=(32, 64, 128, 256)):
    """C4: random hull pairs, centre distance U(0,1.6)(Ra+Rb) → ≈50 % intersecting.  Returns a
    Scene with 2·n_pairs bodies and the pair index arrays (2k, 2k+1)."""
    shapes, radii = hull_library(n_hulls, seed, sizes)
    rng = SplitMix64(seed + 0x1000)
    ha = rng.randint(n_pairs, n_hulls)
    hb = rng.randint(n_pairs, n_hulls)
    qa = rng.quats(n_pairs)
    qb = rng.quats(n_pairs)
    ca = rng.uniform(-10.0, 10.0, n_pairs, 3)
    dirv = rng.quats(n_pairs)[:, :3]
    dirv /= np.maximum(1e-12, np.sqrt((dirv * dirv).sum(axis=1, keepdims=True)))
    dist = rng.uniform(0.0, 1.6, n_pairs) * (radii[ha] + radii[hb])
    cb = ca + dirv * dist[:, None]
    pos = np.empty((2 * n_pairs, 3))
    quat = np.empty((2 * n_pairs, 4))
    sid = np.empty(2 * n_pairs, dtype=np.uint32)
    pos[0::2], pos[1::2] = ca, cb
    quat[0::2], quat[1::2] = qa, qb
    sid[0::2], sid[1::2] = ha, hb
    sc = Scene(shapes, pos, quat, sid)
    return sc, np.arange(0, 2 * n_pairs, 2, dtype=np.uint32), np.arange(1, 2 * n_pairs, 2, dtype=np.uint32)


def scene_c1(side=10, seed=0x5EED0001, spacing=1.2, mesh_boxes=True):
    """C1: static ground box + side³ dynamic unit boxes on a lattice (lowest layer y=1.0, yaw jitter
    u01·0.2 rad).  Body 0 is the ground.  Boxes are 8-vertex hulls, as in physkit::world where
    every body is a mesh::instance (object.h:137)."""
    rng = SplitMix64(seed)
    n = side**3
    ii = np.arange(n)
    gx, gy, gz = ii % side, (ii // side) % side, ii // (side * side)
    off = (side - 1) * spacing / 2.0
    pos = np.stack([gx * spacing - off, 1.0 + gy * spacing, gz * spacing - off], axis=1).astype(np.float64)
    yaw = rng.u01(n) * 0.2
    quat = np.stack([np.zeros(n), np.sin(yaw / 2), np.zeros(n), np.cos(yaw / 2)], axis=1)
    ground_half = (50.0, 0.5, 50.0)
    if mesh_boxes:
        shapes = [("hull", box_vertices(ground_half)), ("hull", box_vertices((0.5, 0.5, 0.5)))]
    else:
        shapes = [("obb", ground_half), ("obb", (0.5, 0.5, 0.5))]
    pos = np.concatenate([[[0.0, -0.5, 0.0]], pos])
    quat = np.concatenate([[list(IDENT)], quat])
    sid = np.concatenate([[0], np.ones(n, dtype=np.uint32)]).astype(np.uint32)
    flags = np.full(n + 1, 2, dtype=np.uint8)
    flags[0] = 3  # static | alive
    return Scene(shapes, pos, quat, sid, flags)


def random_pairs_scene(n_pairs, seed, kinds=("obb", "sphere", "hull", "aabb"), spread=1.2):
    """Differential-test pairs over all shape kinds: body 2k vs 2k+1, centres close enough that
    roughly half intersect."""
    rng = SplitMix64(seed)
    shapes = []
    pos = np.zeros((2 * n_pairs, 3))
    quat = rng.quats(2 * n_pairs)
    kind_pick = rng.randint(2 * n_pairs, len(kinds))
    ca = rng.uniform(-5.0, 5.0, n_pairs, 3)
    off = rng.uniform(-spread, spread, n_pairs, 3)
    pos[0::2] = ca
    pos[1::2] = ca + off
    par = rng.uniform(0.2, 0.8, 2 * n_pairs, 3)
    nv = 4 + rng.randint(2 * n_pairs, 29)
    for i in range(2 * n_pairs):
        k = kinds[kind_pick[i]]
        if k == "obb":
            shapes.append(("obb", par[i]))
        elif k == "sphere":
            shapes.append(("sphere", par[i, 0]))
        elif k == "aabb":
            shapes.append(("aabb", pos[i] - par[i], pos[i] + par[i]))
        else:
            sub = SplitMix64(seed * 7919 + i)
            u = sub.u01(int(nv[i]), 2)
            z = 2.0 * u[:, 0] - 1.0
            t = 2.0 * math.pi * u[:, 1]
            s = np.sqrt(np.maximum(0.0, 1.0 - z * z))
            pts = np.stack([s * np.cos(t), s * np.sin(t), z], axis=1) * par[i]
            shapes.append(("hull", pts))
    sc = Scene(shapes, pos, quat, np.arange(2 * n_pairs))
    return sc, np.arange(0, 2 * n_pairs, 2, dtype=np.uint32), np.arange(1, 2 * n_pairs, 2, dtype=np.uint32)


def _quat_matrix(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def near_touching_scene(n_pairs, seed, kinds=("obb", "sphere", "hull", "bighull", "meshbox", "aabb"), scales=(1e-3, 1.0, 1e3),
                        far=1e4):
    """Pairs (2k, 2k+1) whose gap along a random axis u — B's centre is put at cA + u·(h_A(u) + h_B(−u) + gap) — runs
    through 0, ±1e-12 … ±1e-1 times the pair's size: grazing contacts and grazing misses, which a filter with a margin
    must hand to the exact iteration.  Pairs live up to `far` from the origin at three size scales."""
    rng = SplitMix64(seed)
    gaps = np.concatenate([[0.0], *[[g, -g] for g in 10.0 ** np.arange(-12, 0)]])
    quat = rng.quats(2 * n_pairs)
    kind_pick = rng.randint(2 * n_pairs, len(kinds))
    centre = rng.uniform(-far, far, n_pairs, 3)
    axis = rng.uniform(-1.0, 1.0, n_pairs, 3)
    axis /= np.linalg.norm(axis, axis=1)[:, None]
    axis[::5] = np.eye(3)[np.arange(len(axis[::5])) % 3]  # every fifth pair along a coordinate axis,
    for p in range(1, n_pairs, 3):  # every third along a local axis of A (a face normal, if A is a box: vertex-face contacts)
        axis[p] = _quat_matrix(quat[2 * p])[:, p % 3]
    par = rng.uniform(0.2, 0.8, 2 * n_pairs, 3)
    scale_pick = rng.randint(n_pairs, len(scales))
    shapes, pos = [], np.zeros((2 * n_pairs, 3))
    hval = np.zeros(2 * n_pairs)
    for i in range(2 * n_pairs):
        k = kinds[kind_pick[i]]
        s = scales[scale_pick[i // 2]]
        u = axis[i // 2] if i % 2 == 0 else -axis[i // 2]
        R = _quat_matrix(quat[i])
        if k == "aabb":
            R = np.eye(3)
        l = R.T @ u
        if k in ("obb", "aabb", "meshbox"):
            h = par[i] * s
            hval[i] = np.abs(l) @ h
            if k == "obb":
                shapes.append(("obb", h))
            elif k == "meshbox":
                shapes.append(("hull", np.array([[h[0] if (0x66 >> j) & 1 else -h[0], h[1] if (0xCC >> j) & 1 else -h[1], h[2] if j >= 4 else -h[2]] for j in range(8)])))
            else:
                shapes.append(("aabb", h))  # corners filled in below, once the centre is known
        elif k == "sphere":
            hval[i] = par[i, 0] * s
            shapes.append(("sphere", par[i, 0] * s))
        else:
            sub = SplitMix64(seed * 7919 + i)
            nv = 4 + int(sub.randint(1, 29)[0]) if k == "hull" else 40 + int(sub.randint(1, 160)[0])
            uu = sub.u01(nv, 2)
            z = 2.0 * uu[:, 0] - 1.0
            t = 2.0 * math.pi * uu[:, 1]
            sq = np.sqrt(np.maximum(0.0, 1.0 - z * z))
            pts = np.stack([sq * np.cos(t), sq * np.sin(t), z], axis=1) * par[i] * s
            hval[i] = (pts @ l).max()
            shapes.append(("hull", pts))
    for p in range(n_pairs):
        size = hval[2 * p] + hval[2 * p + 1]
        pos[2 * p] = centre[p]
        pos[2 * p + 1] = centre[p] + axis[p] * (size + gaps[(p * 7 + p // 25) % len(gaps)] * size)
    for i in range(2 * n_pairs):
        if shapes[i][0] == "aabb":
            h = shapes[i][1]
            shapes[i] = ("aabb", pos[i] - h, pos[i] + h)
    sc = Scene(shapes, pos, quat, np.arange(2 * n_pairs))
    return sc, np.arange(0, 2 * n_pairs, 2, dtype=np.uint32), np.arange(1, 2 * n_pairs, 2, dtype=np.uint32)
