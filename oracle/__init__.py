"""CPU ORACLE bindings (test infrastructure — see oracle/pk_oracle.hpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  physkit_b200 never does; its product path fails loudly without the CUDA library.

Shapes are given as a list of specs shared with tests/scenes.py and physkit_b200:
    ("aabb", min3, max3) | ("obb", half3) | ("sphere", r) | ("hull", verts[n,3])
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpk_oracle.so")

KIND_AABB, KIND_OBB, KIND_SPHERE, KIND_HULL = 0, 1, 2, 3

SHAPE_DTYPE = np.dtype(
    [
        ("kind", np.int32),
        ("vert_off", np.uint32),
        ("nverts", np.uint32),
        ("_pad", np.uint32),
        ("a", np.float64, 3),
        ("b", np.float64, 3),
        ("lmin", np.float64, 3),
        ("lmax", np.float64, 3),
    ],
    align=True,
)
assert SHAPE_DTYPE.itemsize == 112


def build(force: bool = False) -> str:
    """Compile oracle/libpk_oracle.so with the committed Makefile (g++ only)."""
    src = [os.path.join(_HERE, f) for f in ("pk_oracle.hpp", "pk_oracle_c.cpp", "Makefile")]
    if (
        force
        or not os.path.exists(_LIB_PATH)
        or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
        L.pko_shape_local_aabb.argtypes = [vp, vp]
        L.pko_bounds.argtypes = [vp, vp, vp, vp, u64, vp]
        L.pko_support.argtypes = [vp, vp, vp, vp, u32, vp, vp]
        L.pko_gjk_epa_pairs.argtypes = [vp, vp, vp, vp, vp, vp, vp, u64, vp, vp, vp, i32]
        L.pko_gjk_epa_pairs.restype = u64
        L.pko_distance_brute_pairs.argtypes = [vp, vp, vp, vp, vp, vp, vp, u64, u32, vp, i32]
        L.pko_distance_brute_pairs.restype = None
        L.pko_contact_points.argtypes = [vp, vp, vp, vp, vp, u64, vp]
        L.pko_man_create.restype = vp
        L.pko_man_destroy.argtypes = [vp]
        L.pko_man_step.argtypes = [vp, vp, vp, vp, u64, vp, vp, vp, vp, vp, vp]
        L.pko_man_step.restype = u64
        L.pko_man_get.argtypes = [vp, vp, vp, vp, u64]
        L.pko_man_get.restype = u64
        L.pko_man_set_impulses.argtypes = [vp, vp]
        L.pko_bvh_create.restype = vp
        L.pko_bvh_destroy.argtypes = [vp]
        L.pko_bvh_add.argtypes = [vp, u32, vp]
        L.pko_bvh_add.restype = u32
        L.pko_bvh_remove.argtypes = [vp, u32]
        L.pko_bvh_update.argtypes = [vp, u32, vp, vp]
        L.pko_bvh_update.restype = i32
        L.pko_bvh_bounds.argtypes = [vp, u32, vp]
        L.pko_bvh_data.argtypes = [vp, u32]
        L.pko_bvh_data.restype = u32
        L.pko_bvh_validate.argtypes = [vp]
        L.pko_bvh_validate.restype = i32
        L.pko_bvh_query.argtypes = [vp, vp, vp, u64, u64]
        L.pko_bvh_query.restype = u64
        L.pko_bp_create.restype = vp
        L.pko_bp_destroy.argtypes = [vp]
        L.pko_bp_add.argtypes = [vp, u32, vp, i32]
        L.pko_bp_remove.argtypes = [vp, u32]
        L.pko_bp_update.argtypes = [vp, u32, vp, vp]
        L.pko_bp_update.restype = i32
        L.pko_bp_calculate.argtypes = [vp]
        L.pko_bp_pairs.argtypes = [vp, vp, u64]
        L.pko_bp_pairs.restype = u64
        L.pko_bp_stored.argtypes = [vp, u32, vp]
        L.pko_world_create.restype = vp
        L.pko_world_destroy.argtypes = [vp]
        L.pko_world_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, u64]
        L.pko_world_step.restype = u64
        L.pko_world_pairs.argtypes = [vp, vp, u64]
        L.pko_world_pairs.restype = u64
        L.pko_world_stored.argtypes = [vp, u32, vp]
        L.pko_query_pairs.argtypes = [vp, u64, vp, u64]
        L.pko_query_pairs.restype = u64
        L.pko_brute_pairs.argtypes = [vp, u64, vp, u64]
        L.pko_brute_pairs.restype = u64
        L.pko_max_threads.restype = i32
        f64 = C.c_double
        L.pko_dynamics_step.argtypes = [u64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, f64, i32, vp]
        L.pko_setup_contacts.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, f64, f64, vp, vp, vp, u64]
        L.pko_setup_contacts.restype = u64
        L.pko_ray_box.argtypes = [vp, vp, vp, f64, vp]
        L.pko_ray_box.restype = i32
        L.pko_bvh_raycast.argtypes = [vp, vp, vp, f64, i32, vp, vp, u64]
        L.pko_bvh_raycast.restype = u64
        L.pko_world_raycast.argtypes = [vp, vp, vp, f64, vp, vp, u64]
        L.pko_world_raycast.restype = u64
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class ShapeTable:
    """Flat shape table + shared vertex pool in the layout of pko_shape_desc."""

    def __init__(self, specs):
        self.specs = list(specs)
        verts = []
        off = 0
        tab = np.zeros(len(self.specs), dtype=SHAPE_DTYPE)
        for i, s in enumerate(self.specs):
            k = s[0]
            if k == "aabb":
                tab[i]["kind"] = KIND_AABB
                tab[i]["a"] = np.asarray(s[1], dtype=np.float64)
                tab[i]["b"] = np.asarray(s[2], dtype=np.float64)
            elif k == "obb":
                tab[i]["kind"] = KIND_OBB
                tab[i]["a"] = np.asarray(s[1], dtype=np.float64)
            elif k == "sphere":
                tab[i]["kind"] = KIND_SPHERE
                tab[i]["a"][0] = float(s[1])
            elif k == "hull":
                v = _f64(s[1]).reshape(-1, 3)
                tab[i]["kind"] = KIND_HULL
                tab[i]["vert_off"] = off
                tab[i]["nverts"] = len(v)
                verts.append(v)
                off += len(v)
            else:
                raise ValueError(k)
        self.verts = np.concatenate(verts) if verts else np.zeros((1, 3))
        self.verts = np.ascontiguousarray(self.verts, dtype=np.float64)
        self.tab = tab
        L = lib()
        for i in range(len(tab)):
            L.pko_shape_local_aabb(C.c_void_p(tab.ctypes.data + i * SHAPE_DTYPE.itemsize), _p(self.verts))

    def local_aabb(self, i):
        return np.concatenate([self.tab[i]["lmin"], self.tab[i]["lmax"]])


def _table(shapes):
    return shapes if isinstance(shapes, ShapeTable) else ShapeTable(shapes)


def bounds(shapes, pos, quat, shape_id):
    """mesh::instance::bounds() per body → [n,6]."""
    t = _table(shapes)
    pos = _f64(pos, (-1, 3))
    quat = _f64(quat, (-1, 4))
    sid = np.ascontiguousarray(shape_id, dtype=np.uint32)
    out = np.empty((len(pos), 6))
    lib().pko_bounds(_p(t.tab), _p(pos), _p(quat), _p(sid), len(pos), _p(out))
    return out


def support(shapes, pos, quat, shape_index, direction):
    t = _table(shapes)
    pos = _f64(pos, (3,))
    quat = _f64(quat, (4,))
    d = _f64(direction, (3,))
    out = np.empty(3)
    lib().pko_support(_p(t.tab), _p(t.verts), _p(pos), _p(quat), int(shape_index), _p(d), _p(out))
    return out


def gjk_epa_pairs(shapes, pos, quat, shape_id, pair_a, pair_b, stats=False, nthreads=1):
    """gjk_epa(body a, body b) per pair → (hit[n] u8, contact[n,10], stats[n,8] | None)."""
    t = _table(shapes)
    pos = _f64(pos, (-1, 3))
    quat = _f64(quat, (-1, 4))
    sid = np.ascontiguousarray(shape_id, dtype=np.uint32)
    pa = np.ascontiguousarray(pair_a, dtype=np.uint32)
    pb = np.ascontiguousarray(pair_b, dtype=np.uint32)
    n = len(pa)
    out = np.empty((n, 10))  # every row is written by the C side
    hit = np.empty(n, dtype=np.uint8)
    st = np.empty((n, 8), dtype=np.int32) if stats else None
    lib().pko_gjk_epa_pairs(
        _p(t.tab), _p(t.verts), _p(pos), _p(quat), _p(sid), _p(pa), _p(pb), n, _p(out), _p(hit), _p(st), int(nthreads)
    )
    return hit, out, st


def distance_brute_pairs(shapes, pos, quat, shape_id, pair_a, pair_b, max_verts=16, nthreads=8):
    """Closest distance of every pair by brute force over all vertex / edge / triangle combinations (pk_oracle.hpp
    brute_distance: the checker of pk_gjk_distance_batch) → d[n]; meaningful for disjoint bodies, NaN above max_verts."""
    t = _table(shapes)
    pos = _f64(pos, (-1, 3))
    quat = _f64(quat, (-1, 4))
    sid = np.ascontiguousarray(shape_id, dtype=np.uint32)
    pa = np.ascontiguousarray(pair_a, dtype=np.uint32)
    pb = np.ascontiguousarray(pair_b, dtype=np.uint32)
    out = np.empty(len(pa))
    lib().pko_distance_brute_pairs(_p(t.tab), _p(t.verts), _p(pos), _p(quat), _p(sid), _p(pa), _p(pb), len(pa), int(max_verts), _p(out), int(nthreads))
    return out


def contact_points(pos, quat, pair_a, pair_b, contacts10):
    """contact_point (collision_phases.h:78-82): witness points of every contact in body-local frames → [n,6]."""
    pos = _f64(pos, (-1, 3))
    quat = _f64(quat, (-1, 4))
    pa = np.ascontiguousarray(pair_a, dtype=np.uint32)
    pb = np.ascontiguousarray(pair_b, dtype=np.uint32)
    c = _f64(contacts10, (-1, 10))
    out = np.empty((len(pa), 6))
    lib().pko_contact_points(_p(pos), _p(quat), _p(pa), _p(pb), _p(c), len(pa), _p(out))
    return out


def gjk_epa(shape_a, pose_a, shape_b, pose_b):
    """One pair from two (spec, (pos, quat_xyzw)) → None | dict(normal, world_a, world_b, depth)."""
    pos = np.array([pose_a[0], pose_b[0]], dtype=np.float64)
    quat = np.array([pose_a[1], pose_b[1]], dtype=np.float64)
    hit, out, _ = gjk_epa_pairs([shape_a, shape_b], pos, quat, [0, 1], [0], [1])
    if not hit[0]:
        return None
    o = out[0]
    return dict(normal=o[0:3].copy(), world_a=o[3:6].copy(), world_b=o[6:9].copy(), depth=float(o[9]))


def query_pairs(boxes6):
    """Static-pose pair set through the faithful dynamic_bvh (add all, query all)."""
    b = _f64(boxes6, (-1, 6))
    n = lib().pko_query_pairs(_p(b), len(b), None, 0)
    out = np.empty(n, dtype=np.uint64)
    lib().pko_query_pairs(_p(b), len(b), _p(out), n)
    return out


def brute_pairs(boxes6):
    b = _f64(boxes6, (-1, 6))
    n = lib().pko_brute_pairs(_p(b), len(b), None, 0)
    out = np.empty(n, dtype=np.uint64)
    lib().pko_brute_pairs(_p(b), len(b), _p(out), n)
    return out


class Dynamics:
    """Rigid-body state of n bodies and the two per-body loops of world::step_impl (src/world.cpp:22-34, 50-55)."""

    def __init__(self, pos, quat, vel, ang_vel, mass, inertia, flags):
        n = len(pos)
        self.pos = _f64(pos, (n, 3)).copy()
        self.quat = _f64(quat, (n, 4)).copy()
        self.vel = _f64(vel, (n, 3)).copy()
        self.ang_vel = _f64(ang_vel, (n, 3)).copy()
        self.acc = np.zeros((n, 3))
        self.torque = np.zeros((n, 3))
        self.mass = _f64(mass, (n,)).copy()
        self.inertia = _f64(inertia, (n, 9)).copy()
        self.flags = np.ascontiguousarray(flags, dtype=np.uint8).copy()

    def _step(self, phase, dt, gravity):
        n = len(self.pos)
        disp = np.zeros((n, 3))
        g = _f64(gravity, (3,))
        lib().pko_dynamics_step(n, _p(self.pos), _p(self.quat), _p(self.vel), _p(self.ang_vel), _p(self.acc), _p(self.torque),
                                _p(self.mass), _p(self.inertia), None, _p(self.flags), _p(g), float(dt), phase, _p(disp))
        return disp

    def integrate_velocities(self, dt, gravity):
        """loop A: returns vel·dt, the displacement handed to broad_phase::update_node."""
        return self._step(0, dt, gravity)

    def integrate_positions(self, dt):
        self._step(1, dt, (0.0, 0.0, 0.0))


def ray_box(origin, direction, box6, max_distance):
    """ray::intersect_distance (bvh.h:59-98): distance or None."""
    out = np.empty(1)
    hit = lib().pko_ray_box(_p(_f64(origin, (3,))), _p(_f64(direction, (3,))), _p(_f64(box6, (6,))), float(max_distance), _p(out))
    return float(out[0]) if hit else None


def max_threads():
    return int(lib().pko_max_threads())


class DynamicBVH:
    """dynamic_bvh (bvh.h:270-535) handle."""

    def __init__(self):
        self.h = lib().pko_bvh_create()

    def __del__(self):
        if getattr(self, "h", None):
            lib().pko_bvh_destroy(self.h)
            self.h = None

    def add(self, obj_id, box6):
        return int(lib().pko_bvh_add(self.h, int(obj_id), _p(_f64(box6, (6,)))))

    def remove_leaf(self, leaf):
        lib().pko_bvh_remove(self.h, int(leaf))

    def update_leaf(self, leaf, box6, disp3):
        return bool(lib().pko_bvh_update(self.h, int(leaf), _p(_f64(box6, (6,))), _p(_f64(disp3, (3,)))))

    def bounds(self, leaf):
        out = np.empty(6)
        lib().pko_bvh_bounds(self.h, int(leaf), _p(out))
        return out

    def data(self, leaf):
        return int(lib().pko_bvh_data(self.h, int(leaf)))

    def validate(self):
        return bool(lib().pko_bvh_validate(self.h))

    def raycast(self, origin, direction, max_distance, closest=False, cap=1 << 16):
        """dynamic_bvh::raycast (bvh.h:346-450): (ids, distances) in traversal order; closest=True drives the
        callback form as a closest-leaf search (the last entry is the closest leaf)."""
        ids = np.empty(cap, dtype=np.uint32)
        d = np.empty(cap, dtype=np.float64)
        n = lib().pko_bvh_raycast(self.h, _p(_f64(origin, (3,))), _p(_f64(direction, (3,))), float(max_distance), int(closest),
                                  _p(ids), _p(d), cap)
        n = min(n, cap)
        return ids[:n].copy(), d[:n].copy()

    def query_aabb(self, box6, stop_after=0, cap=1 << 16):
        out = np.empty(cap, dtype=np.uint32)
        n = lib().pko_bvh_query(self.h, _p(_f64(box6, (6,))), _p(out), cap, int(stop_after))
        return out[: min(n, cap)].copy()


class World:
    """The collision stage of world::step_impl (src/world.cpp:30-46) on flat body arrays."""

    def __init__(self, shapes):
        self.t = _table(shapes)
        self.h = lib().pko_world_create()

    def __del__(self):
        if getattr(self, "h", None):
            lib().pko_world_destroy(self.h)
            self.h = None

    def step(self, pos, quat, disp, shape_id, flags):
        pos = _f64(pos, (-1, 3))
        quat = _f64(quat, (-1, 4))
        disp = _f64(disp, (-1, 3))
        sid = np.ascontiguousarray(shape_id, dtype=np.uint32)
        fl = np.ascontiguousarray(flags, dtype=np.uint8)
        moved = lib().pko_world_step(self.h, _p(self.t.tab), _p(pos), _p(quat), _p(disp), _p(sid), _p(fl), len(pos))
        return int(moved)

    def pairs(self):
        n = lib().pko_world_pairs(self.h, None, 0)
        out = np.empty(n, dtype=np.uint64)
        lib().pko_world_pairs(self.h, _p(out), n)
        return out

    def stored(self, body):
        out = np.empty(6)
        lib().pko_world_stored(self.h, int(body), _p(out))
        return out

    def raycast(self, origin, direction, max_distance, cap=1 << 16):
        """world_base::raycast (core/world.h:260-319): (body ids, distances), merged by distance."""
        ids = np.empty(cap, dtype=np.uint32)
        d = np.empty(cap, dtype=np.float64)
        n = lib().pko_world_raycast(self.h, _p(_f64(origin, (3,))), _p(_f64(direction, (3,))), float(max_distance), _p(ids), _p(d), cap)
        n = min(n, cap)
        return ids[:n].copy(), d[:n].copy()


class Manifolds:
    """narrow_phase's manifold state (collision_phases.h:200-327): step() is narrow_phase::calculate for the
    current pair set given this step's gjk_epa results."""

    def __init__(self):
        self.h = lib().pko_man_create()
        self.count = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().pko_man_destroy(self.h)
            self.h = None

    def step(self, keys, hit, contacts10, pos, quat):
        """→ (began keys, ended keys); keys sorted, hit[n], contacts10[n,10] (rows of hits)."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        hit = np.ascontiguousarray(hit, dtype=np.uint8)
        c = _f64(contacts10, (-1, 10))
        pos = _f64(pos, (-1, 3))
        quat = _f64(quat, (-1, 4))
        n = len(keys)
        began = np.empty(max(n, 1), dtype=np.uint64)
        ended = np.empty(max(n, 1), dtype=np.uint64)
        nb = C.c_uint64()
        ne = C.c_uint64()
        self.count = int(lib().pko_man_step(self.h, _p(keys), _p(hit), _p(c), n, _p(pos), _p(quat), _p(began), C.byref(nb),
                                           _p(ended), C.byref(ne)))
        return np.sort(began[: nb.value]), np.sort(ended[: ne.value])

    def get(self):
        """→ (keys[m], counts[m], points[m,4,13]) in key order; a point is normal(3) local_a(3) local_b(3) depth
        normal_impulse tangent_impulses(2)."""
        m = self.count
        keys = np.empty(max(m, 1), dtype=np.uint64)
        counts = np.empty(max(m, 1), dtype=np.uint32)
        pts = np.zeros((max(m, 1), 4, 13))
        lib().pko_man_get(self.h, _p(keys), _p(counts), _p(pts), max(m, 1))
        return keys[:m], counts[:m], pts[:m]

    def setup_contacts(self, pos, quat, vel, ang_vel, mass, inertia, restitution, friction, dt, gravity_norm):
        """constraint_solver::setup_contacts (constraint.h:1052-1104) → (keys, point index, rows[k, 40])."""
        n = len(pos)
        args = [_f64(pos, (n, 3)), _f64(quat, (n, 4)), _f64(vel, (n, 3)), _f64(ang_vel, (n, 3)), _f64(mass, (n,)), _f64(inertia, (n, 9)),
                _f64(restitution, (n,)), _f64(friction, (n,))]
        cap = 4 * max(1, len(self.get()[0]))
        rows = np.zeros((cap, 40))
        keys = np.zeros(cap, dtype=np.uint64)
        pts = np.zeros(cap, dtype=np.uint32)
        k = lib().pko_setup_contacts(self.h, *[_p(a) for a in args], float(dt), float(gravity_norm), _p(rows), _p(keys), _p(pts), cap)
        return keys[:k].copy(), pts[:k].copy(), rows[:k].copy()

    def set_impulses(self, imp):
        imp = _f64(imp, (-1, 4, 3))
        assert len(imp) == self.count
        lib().pko_man_set_impulses(self.h, _p(imp))
