"""ctypes binding of include/pk_collide.h plus a host-side mirror of the reference's collision API.

Names follow the reference (core/world.h, collision_phases.h, collision.h):
    CollisionWorld.create_rigid / remove_rigid / step   ↔  world_base::create_rigid / remove_rigid / step_impl's ★ calls
    CollisionWorld.pairs()                              ↔  pair_manager::active_pairs() as a sorted list
    CollisionWorld.contacts()                           ↔  one collision_info per colliding manifold
    gjk_epa(a, b)                                       ↔  physkit::gjk_epa (collision.h:61-62)
This module only does plumbing (array marshalling); every computation happens in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("PK_COLLIDE_LIB") or os.path.join(_HERE, "libpk_collide.so")

MODE_WORLD, MODE_QUERY = 0, 1
PK_OK = 0
PK_E_PAIR_OVERFLOW = -5
PK_E_EPA_OVERFLOW = -6
NUM_STAGES = 12

contact_dtype = np.dtype(
    [("key", np.uint64), ("normal", np.float64, 3), ("world_a", np.float64, 3), ("world_b", np.float64, 3), ("depth", np.float64)]
)
assert contact_dtype.itemsize == 88
distance_dtype = np.dtype([("key", np.uint64), ("distance", np.float64), ("point_a", np.float64, 3), ("point_b", np.float64, 3)])
assert distance_dtype.itemsize == 64
solver_row_dtype = np.dtype([("J_v", "<f8", (3,)), ("J_w_a", "<f8", (3,)), ("J_w_b", "<f8", (3,)), ("M_eff", "<f8"), ("bias", "<f8")])
solver_point_dtype = np.dtype([("key", "<u8"), ("manifold", "<u4"), ("point", "<u4"), ("normal", solver_row_dtype),
                               ("tangent1", solver_row_dtype), ("tangent2", solver_row_dtype), ("friction_coeff", "<f8"),
                               ("inv_m_11", "<f8"), ("inv_m_12", "<f8"), ("inv_m_22", "<f8"), ("accumulated", "<f8", (3,))])
assert solver_point_dtype.itemsize == 336
ray_hit_dtype = np.dtype([("ray", np.uint32), ("body", np.uint32), ("distance", np.float64)])
assert ray_hit_dtype.itemsize == 16
RAY_ALL, RAY_CLOSEST = 0, 1


class PkError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"pk_collide error {status}: {msg}")
        self.status = status


class _Gathered(C.Structure):
    _fields_ = [("d_records", C.c_void_p), ("stride_records", C.c_uint64), ("counts", C.POINTER(C.c_uint64)), ("num_ranks", C.c_uint32),
                ("total", C.c_uint64), ("ms", C.c_float)]


class _Config(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("mode", C.c_int32),
        ("max_bodies", C.c_uint32),
        ("max_shapes", C.c_uint32),
        ("max_pairs", C.c_uint64),
        ("max_contacts", C.c_uint64),
        ("max_hull_vertices", C.c_uint64),
        ("num_worlds", C.c_uint32),
        ("shard_rank", C.c_uint32),
        ("shard_count", C.c_uint32),
        ("flags", C.c_uint32),
    ]


class StepResult(C.Structure):
    _fields_ = [
        ("num_pairs", C.c_uint64),
        ("num_contacts", C.c_uint64),
        ("num_moved", C.c_uint64),
        ("pairs_required", C.c_uint64),
        ("epa_overflow", C.c_uint64),
        ("gjk_hits", C.c_uint64),
        ("ms_broadphase", C.c_float),
        ("ms_narrowphase", C.c_float),
        ("ms_total", C.c_float),
        ("step_index", C.c_uint32),
    ]


class _ManifoldResult(C.Structure):
    _fields_ = [("num_manifolds", C.c_uint64), ("num_began", C.c_uint64), ("num_ended", C.c_uint64), ("ms", C.c_float)]


manifold_point_dtype = np.dtype([("normal", "<f8", (3,)), ("local_a", "<f8", (3,)), ("local_b", "<f8", (3,)), ("depth", "<f8"),
                                 ("normal_impulse", "<f8"), ("tangent_impulses", "<f8", (2,))])
manifold_dtype = np.dtype([("key", "<u8"), ("count", "<u4"), ("_pad", "<u4"), ("points", manifold_point_dtype, (4,))])
assert manifold_dtype.itemsize == 432


class _StageTimes(C.Structure):
    _fields_ = [("ms", C.c_float * NUM_STAGES), ("name", C.c_char_p * NUM_STAGES), ("launches", C.c_uint32), ("epa_fallback", C.c_uint32)]


_lib = None

EXPORTS = [
    "pk_abi_version", "pk_selftest_division", "pk_contact_points", "pk_manifolds_enable", "pk_manifolds_update",
    "pk_manifolds", "pk_manifolds_device", "pk_manifold_events", "pk_manifolds_set_impulses", "pk_create", "pk_destroy", "pk_strerror", "pk_last_error",
    "pk_shape_box", "pk_shape_sphere", "pk_shape_hull", "pk_shape_aabb", "pk_shapes_bulk",
    "pk_bodies_resize", "pk_bodies_upload", "pk_bodies_update_pose", "pk_reserve_pairs",
    "pk_comm_get_id", "pk_comm_init", "pk_comm_pose_slice", "pk_comm_allgather_poses", "pk_comm_allgather_contacts",
    "pk_collide_resident", "pk_fetch_results", "pk_collide", "pk_pairs", "pk_contacts",
    "pk_pairs_device", "pk_contacts_device", "pk_stored_bounds", "pk_stage_times_get", "pk_stream",
    "pk_gjk_epa_batch", "pk_gjk_epa_batch_device", "pk_gjk_distance_batch", "pk_gjk_distance_batch_device", "pk_raycast", "pk_raycast_device_ms",
    "pk_dynamics_enable", "pk_dynamics_upload", "pk_dynamics_set_velocities", "pk_dynamics_set_forces",
    "pk_integrate_velocities", "pk_integrate_positions", "pk_dynamics_download", "pk_displacements",
    "pk_material_upload", "pk_contact_rows_setup", "pk_contact_rows", "pk_contact_rows_device",
    "pk_create_multi", "pk_destroy_multi", "pk_multi_size", "pk_multi_ctx", "pk_multi_bodies_resize",
    "pk_multi_bodies_upload", "pk_multi_bodies_update_pose", "pk_multi_collide",
    "pk_device_alloc", "pk_device_free", "pk_memcpy_h2d", "pk_memcpy_d2h", "pk_memcpy_d2d", "pk_host_alloc", "pk_host_free",
]


def library_path() -> str:
    return _LIB_PATH


def load_library():
    """Load libpk_collide.so.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise PkError(-3, f"{_LIB_PATH} is missing: run `python -m physkit_b200.build` (or __graft_entry__.build())")
    L = C.CDLL(_LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    L.pk_strerror.restype = C.c_char_p
    L.pk_strerror.argtypes = [i32]
    L.pk_last_error.restype = C.c_char_p
    L.pk_last_error.argtypes = [vp]
    L.pk_create.argtypes = [vp, vp]
    L.pk_selftest_division.argtypes = [vp, C.c_uint64, C.c_uint64, vp]
    L.pk_contact_points.argtypes = [vp, vp, vp]
    L.pk_raycast.argtypes = [vp, vp, vp, vp, vp, u32, i32, vp, u64, vp]
    L.pk_raycast_device_ms.argtypes = [vp, vp]
    L.pk_dynamics_enable.argtypes = [vp]
    L.pk_dynamics_upload.argtypes = [vp, vp, vp, vp, vp, u32, u32]
    L.pk_dynamics_set_velocities.argtypes = [vp, vp, vp, u32, u32]
    L.pk_dynamics_set_forces.argtypes = [vp, vp, vp, u32, u32]
    L.pk_integrate_velocities.argtypes = [vp, C.c_double, vp]
    L.pk_integrate_positions.argtypes = [vp, C.c_double]
    L.pk_dynamics_download.argtypes = [vp, vp, vp, vp, vp, u32, u32]
    L.pk_displacements.argtypes = [vp, vp, u32, u32]
    L.pk_material_upload.argtypes = [vp, vp, vp, u32, u32]
    L.pk_create_multi.argtypes = [vp, vp, i32, vp]
    L.pk_destroy_multi.argtypes = [vp]
    L.pk_multi_size.argtypes = [vp, vp]
    L.pk_multi_ctx.argtypes = [vp, i32, vp]
    L.pk_multi_bodies_resize.argtypes = [vp, u32]
    L.pk_multi_bodies_upload.argtypes = [vp, vp, vp, vp, vp, vp, vp, u32, u32]
    L.pk_multi_bodies_update_pose.argtypes = [vp, vp, vp, vp, u32, u32]
    L.pk_multi_collide.argtypes = [vp, vp]
    L.pk_contact_rows_setup.argtypes = [vp, C.c_double, C.c_double, vp]
    L.pk_contact_rows.argtypes = [vp, vp, vp]
    L.pk_contact_rows_device.argtypes = [vp, vp, vp, vp]
    L.pk_manifolds_enable.argtypes = [vp, C.c_uint64]
    L.pk_manifolds_update.argtypes = [vp, vp]
    L.pk_manifolds.argtypes = [vp, vp, vp]
    L.pk_manifolds_device.argtypes = [vp, vp, vp]
    L.pk_manifold_events.argtypes = [vp, vp, vp, vp, vp]
    L.pk_manifolds_set_impulses.argtypes = [vp, vp, C.c_uint64]
    L.pk_destroy.argtypes = [vp]
    L.pk_shape_box.argtypes = [vp, vp, vp]
    L.pk_shape_sphere.argtypes = [vp, C.c_double, vp]
    L.pk_shape_hull.argtypes = [vp, vp, u32, vp]
    L.pk_shape_aabb.argtypes = [vp, vp, vp, vp]
    L.pk_shapes_bulk.argtypes = [vp, vp, vp, u32, vp]
    L.pk_bodies_resize.argtypes = [vp, u32]
    L.pk_bodies_upload.argtypes = [vp, vp, vp, vp, vp, vp, vp, u32, u32]
    L.pk_bodies_update_pose.argtypes = [vp, vp, vp, vp, u32, u32]
    L.pk_reserve_pairs.argtypes = [vp, C.c_uint64, C.c_uint64]
    L.pk_comm_get_id.argtypes = [vp]
    L.pk_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.pk_comm_pose_slice.argtypes = [vp, vp, vp]
    L.pk_comm_allgather_poses.argtypes = [vp, C.c_int]
    L.pk_comm_allgather_contacts.argtypes = [vp, vp]
    L.pk_collide_resident.argtypes = [vp, vp]
    L.pk_fetch_results.argtypes = [vp]
    L.pk_collide.argtypes = [vp, vp]
    L.pk_pairs.argtypes = [vp, vp, vp]
    L.pk_contacts.argtypes = [vp, vp, vp]
    L.pk_pairs_device.argtypes = [vp, vp, vp]
    L.pk_contacts_device.argtypes = [vp, vp, vp]
    L.pk_stored_bounds.argtypes = [vp, vp, u32, u32]
    L.pk_stage_times_get.argtypes = [vp, vp]
    L.pk_stream.argtypes = [vp, vp]
    L.pk_gjk_epa_batch.argtypes = [vp, vp, vp, u64, vp, vp]
    L.pk_gjk_epa_batch_device.argtypes = [vp, vp, vp, u64, vp, vp, vp]
    L.pk_gjk_distance_batch.argtypes = [vp, vp, vp, u64, vp, vp]
    L.pk_gjk_distance_batch_device.argtypes = [vp, vp, vp, u64, vp, vp, vp]
    L.pk_device_alloc.argtypes = [vp, C.c_size_t, vp]
    L.pk_device_free.argtypes = [vp, vp]
    L.pk_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    L.pk_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    L.pk_memcpy_d2d.argtypes = [vp, vp, vp, C.c_size_t]
    L.pk_host_alloc.argtypes = [vp, C.c_size_t, vp]
    L.pk_host_free.argtypes = [vp, vp]
    _lib = L
    return L


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _arr(a, dtype, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    return a.reshape(shape) if shape is not None else a


class Context:
    """One pk_ctx (one device, one stream)."""

    def __init__(self, max_bodies, max_pairs, mode=MODE_WORLD, device=0, max_shapes=None, max_contacts=0,
                 max_hull_vertices=0, num_worlds=1, shard_rank=0, shard_count=1, _borrow=None):
        self.L = load_library()
        self._owned = _borrow is None
        if _borrow is not None:  # a context owned by a pk_multi
            self.h = C.c_void_p(_borrow)
            self.mode = mode
            self.result = StepResult()
            self._pinned = []
            return
        cfg = _Config(
            int(device), int(mode), int(max_bodies), int(max_shapes if max_shapes is not None else max(16, max_bodies)),
            int(max_pairs), int(max_contacts), int(max_hull_vertices), int(num_worlds), int(shard_rank), int(shard_count), 0,
        )
        self.h = C.c_void_p()
        self.mode = mode
        st = self.L.pk_create(C.byref(cfg), C.byref(self.h))
        if st != PK_OK:
            self.h = None
            raise PkError(st, self.L.pk_strerror(st).decode())
        self.result = StepResult()
        self._pinned = []

    # -- plumbing
    def _check(self, st, allow=()):
        if st != PK_OK and st not in allow:
            raise PkError(st, f"{self.L.pk_strerror(st).decode()} ({self.L.pk_last_error(self.h).decode()})")
        return st

    def close(self):
        if getattr(self, "h", None):
            for ptr in self._pinned:
                self.L.pk_host_free(self.h, ptr)
            self._pinned = []
            if self._owned:
                self.L.pk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- shapes
    def add_shape(self, spec):
        """spec: ("aabb", min, max) | ("obb", half) | ("sphere", r) | ("hull", verts[n,3]) → shape id."""
        sid = C.c_uint32()
        k = spec[0]
        if k == "obb":
            h = _arr(spec[1], np.float64, (3,))
            self._check(self.L.pk_shape_box(self.h, _p(h), C.byref(sid)))
        elif k == "sphere":
            self._check(self.L.pk_shape_sphere(self.h, float(spec[1]), C.byref(sid)))
        elif k == "hull":
            v = _arr(spec[1], np.float64).reshape(-1, 3)
            self._check(self.L.pk_shape_hull(self.h, _p(v), len(v), C.byref(sid)))
        elif k == "aabb":
            mn, mx = _arr(spec[1], np.float64, (3,)), _arr(spec[2], np.float64, (3,))
            self._check(self.L.pk_shape_aabb(self.h, _p(mn), _p(mx), C.byref(sid)))
        else:
            raise ValueError(k)
        return sid.value

    def add_shapes(self, specs):
        """Add many shapes; boxes/spheres go through one bulk call."""
        specs = list(specs)
        if specs and all(s[0] in ("obb", "sphere") for s in specs):
            kind = np.array([1 if s[0] == "obb" else 2 for s in specs], dtype=np.int32)
            par = np.zeros((len(specs), 3))
            for i, s in enumerate(specs):
                if s[0] == "obb":
                    par[i] = s[1]
                else:
                    par[i, 0] = s[1]
            first = C.c_uint32()
            self._check(self.L.pk_shapes_bulk(self.h, _p(kind), _p(par), len(specs), C.byref(first)))
            return list(range(first.value, first.value + len(specs)))
        return [self.add_shape(s) for s in specs]

    # -- bodies
    def resize(self, n):
        self._check(self.L.pk_bodies_resize(self.h, int(n)))

    def upload(self, pos, quat, disp, shape_id, flags, world_id=None, first=0):
        pos = _arr(pos, np.float64).reshape(-1, 3)
        quat = _arr(quat, np.float64).reshape(-1, 4)
        disp = None if disp is None else _arr(disp, np.float64).reshape(-1, 3)
        sid = _arr(shape_id, np.uint32)
        fl = _arr(flags, np.uint8)
        wid = _arr(world_id, np.uint32)
        self._check(self.L.pk_bodies_upload(self.h, _p(pos), _p(quat), _p(disp), _p(sid), _p(fl), _p(wid), int(first), len(pos)))

    def update_pose(self, pos=None, quat=None, disp=None, first=0, count=None):
        pos = None if pos is None else _arr(pos, np.float64).reshape(-1, 3)
        quat = None if quat is None else _arr(quat, np.float64).reshape(-1, 4)
        disp = None if disp is None else _arr(disp, np.float64).reshape(-1, 3)
        if count is None:
            count = len(pos if pos is not None else (quat if quat is not None else disp))
        self._check(self.L.pk_bodies_update_pose(self.h, _p(pos), _p(quat), _p(disp), int(first), int(count)))

    def pinned_empty(self, shape, dtype):
        """numpy array over page-locked memory (freed with the context)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        ptr = C.c_void_p()
        self._check(self.L.pk_host_alloc(self.h, n, C.byref(ptr)))
        self._pinned.append(ptr)
        buf = (C.c_char * n).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    # -- the stage
    def collide_resident(self, allow_epa_overflow=False):
        allow = (PK_E_EPA_OVERFLOW,) if allow_epa_overflow else ()
        self._check(self.L.pk_collide_resident(self.h, C.byref(self.result)), allow)
        return self.result

    def fetch(self):
        self._check(self.L.pk_fetch_results(self.h))

    def collide(self, allow_epa_overflow=False):
        allow = (PK_E_EPA_OVERFLOW,) if allow_epa_overflow else ()
        self._check(self.L.pk_collide(self.h, C.byref(self.result)), allow)
        return self.result

    # -- one world over several processes (pk_comm_*: NCCL all-gathers inside the library)
    @staticmethod
    def comm_get_id():
        """128 bytes rank 0 hands to the other ranks (numpy uint8 array)."""
        buf = np.zeros(128, dtype=np.uint8)
        st = load_library().pk_comm_get_id(_p(buf))
        if st != PK_OK:
            raise PkError(st, "pk_comm_get_id failed (NCCL not available?)")
        return buf

    def comm_init(self, comm_id, rank, nranks):
        cid = np.ascontiguousarray(comm_id, dtype=np.uint8)
        assert cid.size == 128
        self._check(self.L.pk_comm_init(self.h, _p(cid), int(rank), int(nranks)))

    def comm_pose_slice(self):
        f, c = C.c_uint32(), C.c_uint32()
        self._check(self.L.pk_comm_pose_slice(self.h, C.byref(f), C.byref(c)))
        return int(f.value), int(c.value)

    def comm_allgather_poses(self, pos=True, quat=True, disp=True):
        self._check(self.L.pk_comm_allgather_poses(self.h, (1 if pos else 0) | (2 if quat else 0) | (4 if disp else 0)))

    def comm_allgather_contacts(self):
        """→ (device pointer, stride in records, counts per rank, device ms)."""
        g = _Gathered()
        self._check(self.L.pk_comm_allgather_contacts(self.h, C.byref(g)))
        counts = np.ctypeslib.as_array(g.counts, shape=(g.num_ranks,)).copy() if g.num_ranks else np.zeros(0, np.uint64)
        return g.d_records, int(g.stride_records), counts, float(g.ms)

    def reserve_pairs(self, max_pairs, max_contacts=0):
        """Grow the pair / contact capacities (after PK_E_PAIR_OVERFLOW: reserve, then repeat the step)."""
        self._check(self.L.pk_reserve_pairs(self.h, int(max_pairs), int(max_contacts)))

    def pairs(self):
        ptr, n = C.c_void_p(), C.c_uint64()
        self._check(self.L.pk_pairs(self.h, C.byref(ptr), C.byref(n)))
        if n.value == 0:
            return np.empty(0, dtype=np.uint64)
        buf = (C.c_uint64 * n.value).from_address(ptr.value)
        return np.frombuffer(buf, dtype=np.uint64).copy()

    def contacts(self):
        ptr, n = C.c_void_p(), C.c_uint64()
        self._check(self.L.pk_contacts(self.h, C.byref(ptr), C.byref(n)))
        if n.value == 0:
            return np.empty(0, dtype=contact_dtype)
        buf = (C.c_char * (n.value * 88)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=contact_dtype).copy()

    def contacts_device(self):
        ptr, n = C.c_void_p(), C.c_uint64()
        self._check(self.L.pk_contacts_device(self.h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def pairs_device(self):
        ptr, n = C.c_void_p(), C.c_uint64()
        self._check(self.L.pk_pairs_device(self.h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def stored_bounds(self, first=0, count=None):
        if count is None:
            raise ValueError("count required")
        out = np.empty((count, 6))
        self._check(self.L.pk_stored_bounds(self.h, _p(out), int(first), int(count)))
        return out

    def stage_times(self):
        st = _StageTimes()
        self._check(self.L.pk_stage_times_get(self.h, C.byref(st)))
        d = {st.name[k].decode(): float(st.ms[k]) for k in range(NUM_STAGES)}
        self.epa_fallback = int(st.epa_fallback)
        return d, int(st.launches)

    def contact_points(self):
        """Body-local witness points [n, 6] (local_a, local_b) of the last step's contacts."""
        p = C.c_void_p()
        n = C.c_uint64()
        self._check(self.L.pk_contact_points(self.h, C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros((0, 6))
        buf = (C.c_double * (6 * n.value)).from_address(p.value)
        return np.frombuffer(buf, dtype=np.float64).reshape(-1, 6).copy()

    # -- integrator: the two per-body loops of world::step_impl on the device (src/world.cpp:22-34, 50-55)
    def dynamics_enable(self):
        self._check(self.L.pk_dynamics_enable(self.h))

    def dynamics_upload(self, vel, ang_vel, mass, inertia_local, first=0):
        v = _arr(vel, np.float64).reshape(-1, 3)
        w = _arr(ang_vel, np.float64).reshape(-1, 3)
        m = _arr(mass, np.float64).reshape(-1)
        it = _arr(inertia_local, np.float64).reshape(-1, 9)
        assert len(v) == len(w) == len(m) == len(it)
        self._check(self.L.pk_dynamics_upload(self.h, _p(v), _p(w), _p(m), _p(it), int(first), len(v)))

    def dynamics_set_velocities(self, vel, ang_vel, first=0):
        v = _arr(vel, np.float64).reshape(-1, 3)
        w = _arr(ang_vel, np.float64).reshape(-1, 3)
        self._check(self.L.pk_dynamics_set_velocities(self.h, _p(v), _p(w), int(first), len(v)))

    def dynamics_set_forces(self, acc, torque, first=0):
        a = _arr(acc, np.float64).reshape(-1, 3)
        t = _arr(torque, np.float64).reshape(-1, 3)
        self._check(self.L.pk_dynamics_set_forces(self.h, _p(a), _p(t), int(first), len(a)))

    def integrate_velocities(self, dt, gravity=(0.0, -9.81, 0.0)):
        g = np.ascontiguousarray(gravity, dtype=np.float64)
        self._check(self.L.pk_integrate_velocities(self.h, float(dt), _p(g)))

    def integrate_positions(self, dt):
        self._check(self.L.pk_integrate_positions(self.h, float(dt)))

    def dynamics_download(self, count, first=0):
        """(pos, quat, vel, ang_vel) of bodies [first, first + count)."""
        pos, quat = np.empty((count, 3)), np.empty((count, 4))
        vel, w = np.empty((count, 3)), np.empty((count, 3))
        self._check(self.L.pk_dynamics_download(self.h, _p(pos), _p(quat), _p(vel), _p(w), int(first), int(count)))
        return pos, quat, vel, w

    def displacements(self, count, first=0):
        d = np.empty((count, 3))
        self._check(self.L.pk_displacements(self.h, _p(d), int(first), int(count)))
        return d

    # -- contact rows: constraint_solver::setup_contacts on the device (constraint.h:1052-1104)
    def material_upload(self, restitution, friction, first=0):
        r = _arr(restitution, np.float64).reshape(-1)
        f = _arr(friction, np.float64).reshape(-1)
        self._check(self.L.pk_material_upload(self.h, _p(r), _p(f), int(first), len(r)))

    def contact_rows_setup(self, dt, gravity_norm=9.81):
        n = C.c_uint64()
        self._check(self.L.pk_contact_rows_setup(self.h, float(dt), float(gravity_norm), C.byref(n)))
        return int(n.value)

    def contact_rows(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self.L.pk_contact_rows(self.h, C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=solver_point_dtype)
        buf = (C.c_char * (solver_point_dtype.itemsize * n.value)).from_address(p.value)
        return np.frombuffer(buf, dtype=solver_point_dtype).copy()

    def contact_rows_device_ms(self):
        p, n, ms = C.c_void_p(), C.c_uint64(), C.c_float()
        self._check(self.L.pk_contact_rows_device(self.h, C.byref(p), C.byref(n), C.byref(ms)))
        return float(ms.value)

    # -- ray casts over the tree of the last step (world_base::raycast, core/world.h:260-319)
    def raycast(self, origins, directions, max_distance, world=None, mode=RAY_ALL, capacity=None):
        """Batch of rays → structured array (ray, body, distance) sorted by (ray, body).  Grows the result
        buffer and retries once when the library reports the capacity it needs."""
        o = np.ascontiguousarray(origins, dtype=np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(directions, dtype=np.float64).reshape(-1, 3)
        n = len(o)
        assert len(d) == n
        md = np.ascontiguousarray(np.broadcast_to(np.asarray(max_distance, dtype=np.float64), (n,)))
        w = None if world is None else np.ascontiguousarray(world, dtype=np.uint32)
        cap = int(capacity) if capacity is not None else max(4 * n, 1024)
        for _ in range(2):
            out = np.empty(cap, dtype=ray_hit_dtype)
            found = C.c_uint64()
            st = self.L.pk_raycast(self.h, _p(o), _p(d), _p(md), _p(w), n, int(mode), _p(out), cap, C.byref(found))
            if st == PK_E_PAIR_OVERFLOW and capacity is None:
                cap = int(found.value)
                continue
            self._check(st)
            return out[: found.value].copy()
        self._check(st)

    def raycast_device_ms(self):
        ms = C.c_float()
        self._check(self.L.pk_raycast_device_ms(self.h, C.byref(ms)))
        return float(ms.value)

    # -- manifolds (narrow_phase state on the device)
    def manifolds_enable(self, capacity):
        self._check(self.L.pk_manifolds_enable(self.h, C.c_uint64(int(capacity))))

    def manifolds_update(self):
        """→ (num_manifolds, num_began, num_ended, device ms) for the step just computed."""
        r = _ManifoldResult()
        self._check(self.L.pk_manifolds_update(self.h, C.byref(r)))
        return int(r.num_manifolds), int(r.num_began), int(r.num_ended), float(r.ms)

    def manifolds(self):
        """Structured array (manifold_dtype), sorted by key."""
        p = C.c_void_p()
        n = C.c_uint64()
        self._check(self.L.pk_manifolds(self.h, C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=manifold_dtype)
        buf = (C.c_char * (manifold_dtype.itemsize * n.value)).from_address(p.value)
        return np.frombuffer(buf, dtype=manifold_dtype).copy()

    def manifold_events(self):
        """→ (began keys, ended keys), each sorted."""
        pb, pe = C.c_void_p(), C.c_void_p()
        nb, ne = C.c_uint64(), C.c_uint64()
        self._check(self.L.pk_manifold_events(self.h, C.byref(pb), C.byref(nb), C.byref(pe), C.byref(ne)))

        def take(p, n):
            if n.value == 0:
                return np.zeros(0, dtype=np.uint64)
            return np.frombuffer((C.c_uint64 * n.value).from_address(p.value), dtype=np.uint64).copy()

        return take(pb, nb), take(pe, ne)

    def manifolds_set_impulses(self, imp):
        imp = np.ascontiguousarray(imp, dtype=np.float64).reshape(-1, 4, 3)
        self._check(self.L.pk_manifolds_set_impulses(self.h, _p(imp), C.c_uint64(len(imp))))

    def selftest_division(self, seed, samples):
        bad = C.c_uint64()
        self._check(self.L.pk_selftest_division(self.h, C.c_uint64(seed), C.c_uint64(samples), C.byref(bad)))
        return int(bad.value)

    def stream(self):
        s = C.c_void_p()
        self._check(self.L.pk_stream(self.h, C.byref(s)))
        return s.value or 0

    # -- narrowphase only
    def gjk_epa_batch(self, pair_a, pair_b, allow_epa_overflow=False):
        pa = _arr(pair_a, np.uint32)
        pb = _arr(pair_b, np.uint32)
        n = len(pa)
        out = np.zeros(n, dtype=contact_dtype)
        hit = np.zeros(n, dtype=np.uint8)
        allow = (PK_E_EPA_OVERFLOW,) if allow_epa_overflow else ()
        self._check(self.L.pk_gjk_epa_batch(self.h, _p(pa), _p(pb), n, _p(out), _p(hit)), allow)
        return hit, out

    def gjk_distance_batch(self, pair_a, pair_b):
        """→ (separated[n] u8, records[n] distance_dtype): closest distance and closest points of every pair
        (pk_gjk_distance_batch; separated = 0, distance 0 for pairs that touch or overlap)."""
        pa = _arr(pair_a, np.uint32)
        pb = _arr(pair_b, np.uint32)
        n = len(pa)
        out = np.zeros(n, dtype=distance_dtype)
        sep = np.zeros(n, dtype=np.uint8)
        self._check(self.L.pk_gjk_distance_batch(self.h, _p(pa), _p(pb), n, _p(out), _p(sep)))
        return sep, out

    def gjk_distance_batch_device(self, d_a, d_b, n, d_out, d_sep):
        ms = C.c_float()
        self._check(self.L.pk_gjk_distance_batch_device(self.h, C.c_void_p(d_a), C.c_void_p(d_b), int(n), C.c_void_p(d_out), C.c_void_p(d_sep), C.byref(ms)))
        return float(ms.value)

    def device_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(self.L.pk_device_alloc(self.h, int(nbytes), C.byref(p)))
        return p.value

    def device_free(self, ptr):
        self._check(self.L.pk_device_free(self.h, C.c_void_p(ptr)))

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self._check(self.L.pk_memcpy_h2d(self.h, C.c_void_p(dptr), _p(arr), arr.nbytes))

    def d2h(self, arr, dptr):
        self._check(self.L.pk_memcpy_d2h(self.h, _p(arr), C.c_void_p(dptr), arr.nbytes))

    def d2d(self, dst_ptr, src_ptr, nbytes):
        self._check(self.L.pk_memcpy_d2d(self.h, C.c_void_p(dst_ptr), C.c_void_p(src_ptr), int(nbytes)))

    def gjk_epa_batch_device(self, d_a, d_b, n, d_out, d_hit, allow_epa_overflow=False):
        ms = C.c_float()
        allow = (PK_E_EPA_OVERFLOW,) if allow_epa_overflow else ()
        self._check(
            self.L.pk_gjk_epa_batch_device(self.h, C.c_void_p(d_a), C.c_void_p(d_b), int(n), C.c_void_p(d_out), C.c_void_p(d_hit), C.byref(ms)),
            allow,
        )
        return float(ms.value)


class MultiContext:
    """pk_multi: one world sharded over several devices from this process (pk_create_multi).  `ctx[i]` are
    Context views of the per-device contexts; shapes go to every one of them, bodies through upload()."""

    def __init__(self, devices, max_bodies, max_pairs, mode=MODE_WORLD, max_shapes=None, max_contacts=0, max_hull_vertices=0, num_worlds=1):
        self.L = load_library()
        cfg = _Config(0, int(mode), int(max_bodies), int(max_shapes if max_shapes is not None else max(16, max_bodies)), int(max_pairs),
                      int(max_contacts), int(max_hull_vertices), int(num_worlds), 0, 1, 0)
        dev = (C.c_int * len(devices))(*[int(d) for d in devices])
        self.h = C.c_void_p()
        st = self.L.pk_create_multi(C.byref(cfg), dev, len(devices), C.byref(self.h))
        if st != PK_OK:
            self.h = None
            raise PkError(st, self.L.pk_strerror(st).decode())
        n = C.c_int()
        self.L.pk_multi_size(self.h, C.byref(n))
        self.ctx = []
        for i in range(n.value):
            p = C.c_void_p()
            self.L.pk_multi_ctx(self.h, i, C.byref(p))
            self.ctx.append(Context(max_bodies, max_pairs, mode=mode, _borrow=p.value))
        self.result = StepResult()

    def _check(self, st):
        if st != PK_OK:
            raise PkError(st, self.L.pk_strerror(st).decode())

    def add_shapes(self, specs):
        ids = [c.add_shapes(specs) for c in self.ctx]
        assert all(i == ids[0] for i in ids)
        return ids[0]

    def resize(self, n):
        self._check(self.L.pk_multi_bodies_resize(self.h, int(n)))

    def upload(self, pos, quat, disp, shape_id, flags, world_id=None, first=0):
        pos = _arr(pos, np.float64).reshape(-1, 3)
        quat = _arr(quat, np.float64).reshape(-1, 4)
        disp = None if disp is None else _arr(disp, np.float64).reshape(-1, 3)
        sid = _arr(shape_id, np.uint32)
        fl = _arr(flags, np.uint8)
        wid = _arr(world_id, np.uint32)
        self._check(self.L.pk_multi_bodies_upload(self.h, _p(pos), _p(quat), _p(disp), _p(sid), _p(fl), _p(wid), int(first), len(pos)))

    def collide(self):
        self._check(self.L.pk_multi_collide(self.h, C.byref(self.result)))
        return self.result

    def close(self):
        if getattr(self, "h", None):
            for c in self.ctx:
                c.close()
            self.L.pk_destroy_multi(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CollisionWorld:
    """Host mirror of the collision stage of physkit::world (core/world.h:202-242, src/world.cpp:30-46).

    Bodies are created with create_rigid (id = slot index, reused after remove_rigid like the reference
    arena, detail/arena.h:68-87); step(pos, quat, disp) runs update_node for every dynamic body,
    calculate_pairs and gjk_epa per active pair on the GPU and returns the step counters.
    """

    def __init__(self, max_bodies, max_pairs, device=0, mode=MODE_WORLD, max_hull_vertices=1 << 16, **kw):
        self.ctx = Context(max_bodies, max_pairs, mode=mode, device=device, max_hull_vertices=max_hull_vertices, **kw)
        self.cap = max_bodies
        self.pos = np.zeros((max_bodies, 3))
        self.quat = np.tile(np.array([0.0, 0.0, 0.0, 1.0]), (max_bodies, 1))
        self.disp = np.zeros((max_bodies, 3))
        self.shape_id = np.zeros(max_bodies, dtype=np.uint32)
        self.flags = np.zeros(max_bodies, dtype=np.uint8)
        self.n = 0
        self._free = []
        self._dirty = True

    def add_shape(self, spec):
        return self.ctx.add_shape(spec)

    def create_rigid(self, shape_id, pos, quat=(0.0, 0.0, 0.0, 1.0), is_static=False):
        if self._free:
            i = self._free.pop()
        else:
            if self.n >= self.cap:
                raise PkError(-1, "max_bodies exceeded")
            i = self.n
            self.n += 1
        self.pos[i] = pos
        self.quat[i] = quat
        self.disp[i] = 0.0
        self.shape_id[i] = shape_id
        self.flags[i] = 2 | (1 if is_static else 0)
        self._dirty = True
        return i

    def remove_rigid(self, i):
        self.flags[i] = 0
        self._free.append(i)
        self._dirty = True

    def set_pose(self, i, pos, quat=None, disp=None):
        self.pos[i] = pos
        if quat is not None:
            self.quat[i] = quat
        if disp is not None:
            self.disp[i] = disp

    def step(self, fetch=True):
        n = self.n
        self.ctx.resize(n)
        if self._dirty:
            self.ctx.upload(self.pos[:n], self.quat[:n], self.disp[:n], self.shape_id[:n], self.flags[:n])
            self._dirty = False
        else:
            self.ctx.update_pose(self.pos[:n], self.quat[:n], self.disp[:n])
        return self.ctx.collide() if fetch else self.ctx.collide_resident()

    def pairs(self):
        return self.ctx.pairs()

    def contacts(self):
        return self.ctx.contacts()

    def close(self):
        self.ctx.close()


def gjk_epa(shape_a, pose_a, shape_b, pose_b, device=0):
    """physkit::gjk_epa(a, b) for one pair: (spec, (pos, quat_xyzw)) ×2 → None | dict."""
    with Context(2, 4, mode=MODE_QUERY, device=device, max_shapes=4, max_hull_vertices=4096) as ctx:
        ia, ib = ctx.add_shape(shape_a), ctx.add_shape(shape_b)
        ctx.resize(2)
        ctx.upload([pose_a[0], pose_b[0]], [pose_a[1], pose_b[1]], None, [ia, ib], [2, 2])
        hit, out = ctx.gjk_epa_batch([0], [1])
        if not hit[0]:
            return None
        o = out[0]
        return dict(normal=o["normal"].copy(), world_a=o["world_a"].copy(), world_b=o["world_b"].copy(), depth=float(o["depth"]))
