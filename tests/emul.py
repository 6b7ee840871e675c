"""Host-side run of the library's EPA kernels (tests/cpp/epa_emul.cpp through tests/cpp/simt_host.h): TEST
INFRASTRUCTURE for the container without a GPU.  The kernels' own source is compiled by g++, one OS thread per CUDA
thread, warp votes / shuffles as barriers; a test compares the records with the oracle.  The product never loads this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "epa_emul.cpp")
OUT = os.path.join(ROOT, "tests", "cpp", "_build", "libepa_emul.so")
CUDA_INC = "/usr/local/cuda/include"

SHAPE_REC = np.dtype([("a", "<f8", 3), ("b", "<f8", 3), ("kind", "<i4"), ("vert_off", "<u4"), ("nverts", "<u4"), ("mesh_box", "<u4")])
CONTACT = np.dtype([("key", "<u8"), ("normal", "<f8", 3), ("world_a", "<f8", 3), ("world_b", "<f8", 3), ("depth", "<f8")])
DISTANCE = np.dtype([("key", "<u8"), ("distance", "<f8"), ("point_a", "<f8", 3), ("point_b", "<f8", 3)])
assert SHAPE_REC.itemsize == 64 and CONTACT.itemsize == 88 and DISTANCE.itemsize == 64

KIND = {"aabb": 0, "obb": 1, "sphere": 2, "hull": 3}


def available() -> bool:
    return os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h"))


def build(force: bool = False) -> str:
    deps = [SRC, os.path.join(ROOT, "tests", "cpp", "simt_host.h")] + [
        os.path.join(ROOT, "physkit_b200", "csrc", f) for f in ("pk_common.cuh", "pk_narrowphase.cuh", "pk_epa_coop.cuh", "pk_gjk_filter.cuh", "pk_distance.cuh", "pk_sort.cuh")
    ]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O1", "-g", "-std=c++20", "-ffp-contract=off", "-pthread", "-fPIC", "-shared", "-I" + CUDA_INC, SRC, "-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stderr[-4000:])
    return OUT


def shape_table(specs):
    """ShapeRec table + vertex pool exactly as pk_shape_box / _sphere / _aabb / _hull (pk_api.cu) build them."""
    tab = np.zeros(len(specs), dtype=SHAPE_REC)
    verts = []
    off = 0
    for i, s in enumerate(specs):
        k = s[0]
        tab[i]["kind"] = KIND[k]
        if k == "aabb":
            tab[i]["a"] = np.asarray(s[1], dtype=np.float64)
            tab[i]["b"] = np.asarray(s[2], dtype=np.float64)
        elif k == "obb":
            tab[i]["a"] = np.asarray(s[1], dtype=np.float64)
        elif k == "sphere":
            tab[i]["a"][0] = float(s[1])
        else:
            v = np.ascontiguousarray(s[1], dtype=np.float64).reshape(-1, 3)
            if off & 1:
                verts.append(np.zeros((1, 3)))
                off += 1
            tab[i]["vert_off"] = off
            tab[i]["nverts"] = len(v)
            tab[i]["a"] = v.min(axis=0)
            tab[i]["b"] = v.max(axis=0)
            hb = tab[i]["b"]
            if len(v) == 8 and (hb > 0).all():
                want = np.array([[hb[0] if (0x66 >> j) & 1 else -hb[0], hb[1] if (0xCC >> j) & 1 else -hb[1], hb[2] if j >= 4 else -hb[2]] for j in range(8)])
                tab[i]["mesh_box"] = 1 if np.array_equal(v, want) else 0
            verts.append(v)
            off += len(v)
    pool = np.ascontiguousarray(np.concatenate(verts) if verts else np.zeros((2, 3)), dtype=np.float64)
    return tab, pool


def gjk_epa_pairs(specs, pos, quat, shape_id, pair_a, pair_b, capacity=None, mirror=False, arrival=2, nblocks=1):
    """→ (hit[n] u8, contacts[n] CONTACT, stats dict), as pk_gjk_epa_batch delivers them."""
    lib = C.CDLL(build())
    tab, pool = shape_table(specs)
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
    quat = np.ascontiguousarray(quat, dtype=np.float64).reshape(-1, 4)
    sid = np.ascontiguousarray(shape_id, dtype=np.uint32)
    pa = np.ascontiguousarray(pair_a, dtype=np.uint32)
    pb = np.ascontiguousarray(pair_b, dtype=np.uint32)
    n = len(pa)
    cap = n if capacity is None else int(capacity)
    out = np.zeros(n, dtype=CONTACT)
    hit = np.zeros(n, dtype=np.uint8)
    stats = np.zeros(8, dtype=np.uint64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.emu_gjk_epa(p(tab), C.c_uint64(len(tab)), p(pool), C.c_uint64(len(pool)), p(pos), p(quat), p(sid), p(pa), p(pb), C.c_uint64(n), C.c_uint64(cap),
                         p(out), p(hit), C.c_int(1 if mirror else 0), C.c_int(arrival), C.c_int(nblocks), p(stats))
    if rc != 0:
        raise RuntimeError(f"emu_gjk_epa failed: {rc}")
    names = ["gjk_hits", "restarted_in_heap_mode", "handed_to_epa_kernel", "valid", "dropped", "class0", "class1", "class2"]
    return hit, out, dict(zip(names, (int(x) for x in stats)))


def filter_pairs(specs, pos, quat, shape_id, pair_a, pair_b, iters=2):
    """→ dropped[n] u8: 1 where gjk_filter_kernel's per-pair decision (certainly_separated, FP32) is "separated"."""
    lib = C.CDLL(build())
    tab, pool = shape_table(specs)
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
    quat = np.ascontiguousarray(quat, dtype=np.float64).reshape(-1, 4)
    sid = np.ascontiguousarray(shape_id, dtype=np.uint32)
    pa = np.ascontiguousarray(pair_a, dtype=np.uint32)
    pb = np.ascontiguousarray(pair_b, dtype=np.uint32)
    out = np.zeros(len(pa), dtype=np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.emu_filter(p(tab), C.c_uint64(len(tab)), p(pool), C.c_uint64(len(pool)), p(pos), p(quat), p(sid), p(pa), p(pb), C.c_uint64(len(pa)),
                        C.c_int(iters), p(out))
    if rc != 0:
        raise RuntimeError(f"emu_filter failed: {rc}")
    return out


def distance_pairs(specs, pos, quat, shape_id, pair_a, pair_b):
    """→ (separated[n] u8, records[n] DISTANCE): gjk_distance_pair (pk_distance.cuh) per pair, the source
    gjk_distance_kernel runs on the device."""
    lib = C.CDLL(build())
    tab, pool = shape_table(specs)
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
    quat = np.ascontiguousarray(quat, dtype=np.float64).reshape(-1, 4)
    sid = np.ascontiguousarray(shape_id, dtype=np.uint32)
    pa = np.ascontiguousarray(pair_a, dtype=np.uint32)
    pb = np.ascontiguousarray(pair_b, dtype=np.uint32)
    out = np.zeros(len(pa), dtype=DISTANCE)
    sep = np.zeros(len(pa), dtype=np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.emu_distance(p(tab), C.c_uint64(len(tab)), p(pool), C.c_uint64(len(pool)), p(pos), p(quat), p(sid), p(pa), p(pb), C.c_uint64(len(pa)),
                          p(out), p(sep))
    if rc != 0:
        raise RuntimeError(f"emu_distance failed: {rc}")
    return sep, out


# ---- the radix sort and the flag scan (tests/cpp/sort_emul.cpp)
SORT_SRC = os.path.join(ROOT, "tests", "cpp", "sort_emul.cpp")
SORT_OUT = os.path.join(ROOT, "tests", "cpp", "_build", "libsort_emul.so")


def build_sort(force: bool = False) -> str:
    deps = [SORT_SRC, os.path.join(ROOT, "tests", "cpp", "simt_host.h")] + [os.path.join(ROOT, "physkit_b200", "csrc", f) for f in ("pk_common.cuh", "pk_sort.cuh")]
    if not force and os.path.exists(SORT_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(SORT_OUT) for d in deps):
        return SORT_OUT
    os.makedirs(os.path.dirname(SORT_OUT), exist_ok=True)
    cmd = ["g++", "-O1", "-g", "-std=c++20", "-ffp-contract=off", "-pthread", "-fPIC", "-shared", "-I" + CUDA_INC, SORT_SRC, "-o", SORT_OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stderr[-4000:])
    return SORT_OUT


def radix_sort(keys, vals, shifts, lowbits=0, tile=True, n_dev=None):
    """→ (keys, vals) sorted by pk_sort.cuh's kernels over the listed byte shifts (vals may be None).  tile: all passes in
    one launch (radix_sort_tile_kernel, n ≤ 4096), else hist / scan / scatter per pass.  n_dev: the element count as a
    device-side counter smaller than len(keys)."""
    lib = C.CDLL(build_sort())
    k = np.ascontiguousarray(keys, dtype=np.uint64).copy()
    v = None if vals is None else np.ascontiguousarray(vals, dtype=np.uint32).copy()
    sh = np.ascontiguousarray(shifts, dtype=np.int32)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    rc = lib.emu_radix_sort(p(k), p(v), C.c_uint64(len(k)), C.c_uint64(len(k) + 1 if n_dev is None else int(n_dev)), p(sh), C.c_int(len(sh)),
                            C.c_int(lowbits), C.c_int(1 if tile else 0))
    if rc != 0:
        raise RuntimeError(f"emu_radix_sort failed: {rc}")
    return k, v


def flag_scan(flags, n_dev=None):
    """→ (exclusive scan u32[n], total) by flag_tile_sum / tile_sum_scan / flag_scan_apply (the contact slots of the hits)."""
    lib = C.CDLL(build_sort())
    f = np.ascontiguousarray(flags, dtype=np.uint8)
    out = np.full(len(f), 0xFFFFFFFF, dtype=np.uint32)
    total = C.c_ulonglong(0)
    rc = lib.emu_flag_scan(f.ctypes.data_as(C.c_void_p), C.c_uint64(len(f)), C.c_uint64(len(f) + 1 if n_dev is None else int(n_dev)),
                           out.ctypes.data_as(C.c_void_p), C.byref(total))
    if rc != 0:
        raise RuntimeError(f"emu_flag_scan failed: {rc}")
    return out, int(total.value)


# ---- the broadphase (tests/cpp/broad_emul.cpp)
BROAD_SRC = os.path.join(ROOT, "tests", "cpp", "broad_emul.cpp")
BROAD_OUT = os.path.join(ROOT, "tests", "cpp", "_build", "libbroad_emul.so")


def build_broad(force: bool = False) -> str:
    deps = [BROAD_SRC, os.path.join(ROOT, "tests", "cpp", "simt_host.h")] + [
        os.path.join(ROOT, "physkit_b200", "csrc", f) for f in ("pk_common.cuh", "pk_sort.cuh", "pk_broadphase.cuh")]
    if not force and os.path.exists(BROAD_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(BROAD_OUT) for d in deps):
        return BROAD_OUT
    os.makedirs(os.path.dirname(BROAD_OUT), exist_ok=True)
    cmd = ["g++", "-O1", "-g", "-std=c++20", "-ffp-contract=off", "-pthread", "-fPIC", "-shared", "-I" + CUDA_INC, BROAD_SRC, "-o", BROAD_OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stderr[-4000:])
    return BROAD_OUT


class BroadWorld:
    """The library's broadphase kernels stepped on the host in the order pk_collide_resident launches them: the
    counterpart of oracle.World (same step() / pairs() / stored())."""

    def __init__(self, specs, max_bodies, mode_query=False, rows=True, num_worlds=1, shard=(0, 1)):
        self.lib = C.CDLL(build_broad())
        self.lib.emu_broad_create.restype = C.c_void_p
        self.lib.emu_broad_step.restype = C.c_longlong
        self.tab, _ = shape_table(specs)
        self.h = C.c_void_p(self.lib.emu_broad_create(C.c_uint32(max_bodies)))
        self.mode_query, self.rows, self.num_worlds, self.shard = mode_query, rows, num_worlds, shard
        self.n = 0
        self.npairs = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.emu_broad_destroy(self.h)
            self.h = None

    def step(self, pos, quat, disp, shape_id, flags, world_id=None):
        """→ number of bodies re-inserted (moved); raises on a full pair row (the library repeats such a step in list form)."""
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        quat = np.ascontiguousarray(quat, dtype=np.float64).reshape(-1, 4)
        disp = np.ascontiguousarray(disp, dtype=np.float64).reshape(-1, 3)
        sid = np.ascontiguousarray(shape_id, dtype=np.uint32)
        fl = np.ascontiguousarray(flags, dtype=np.uint8)
        wid = None if world_id is None else np.ascontiguousarray(world_id, dtype=np.uint32)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        self.n = len(pos)
        r = self.lib.emu_broad_step(self.h, p(self.tab), p(pos), p(quat), p(disp), p(sid), p(fl), p(wid), C.c_uint32(self.num_worlds), C.c_uint32(self.n),
                                    C.c_int(1 if self.mode_query else 0), C.c_int(1 if self.rows else 0), C.c_uint32(self.shard[0]), C.c_uint32(self.shard[1]))
        if r < 0:
            raise RuntimeError(f"emu_broad_step: {r}" + (" (a pair row overflowed)" if r == -1 else ""))
        self.npairs = int(r)
        moved = C.c_ulonglong(0)
        self.lib.emu_broad_get(self.h, None, None, C.c_uint32(0), C.byref(moved))
        return int(moved.value)

    def pairs(self):
        out = np.empty(self.npairs, dtype=np.uint64)
        self.lib.emu_broad_get(self.h, out.ctypes.data_as(C.c_void_p), None, C.c_uint32(0), None)
        return out

    def stored(self):
        out = np.empty((self.n, 6))
        self.lib.emu_broad_get(self.h, None, out.ctypes.data_as(C.c_void_p), C.c_uint32(self.n), None)
        return out
