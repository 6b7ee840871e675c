#!/usr/bin/env python
"""Regenerate the golden vectors of this directory:  python tests/golden/make_golden.py

The reference itself cannot be built in this image (C++26 + mp-units + Eigen 5 + abseil), so
these vectors come from the CPU oracle (oracle/, pinned on the reference's own known-answer tests, see
tests/kat_cases.py) on seeded scenes of tests/scenes.py.  They freeze today's agreed results bit for bit:
tests/test_golden.py checks that the oracle still reproduces them (CPU) and that the CUDA library does
(GPU), so a change of either side that moves a single bit shows up even if both sides move together.

  world_c3_side6.npz   3 steps of the C3 generator (216 spheres/boxes, world-mode fat AABBs): pair keys per
                       step, GJK hit flags and EPA contacts of the last step
  pairs_mixed.npz      400 random pairs over obb / sphere / hull / aabb shapes: hit flags and contacts
  rows_c3_side8.npz    the rows around the stage on 4 steps of the C3 generator at side 8: world ray
                       casts of 64 seeded rays, manifold keys / counts, contact rows of the last step, and the
                       integrator's velocities / positions after 3 steps of loops A + B (orientations are not
                       frozen: they go through libm's sin / cos)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle  # noqa: E402
from scenes import random_pairs_scene, scene_c3  # noqa: E402


def world_c3(side=6, steps=3):
    sc = scene_c3(side=side)
    w = oracle.World(sc.shapes)
    pos = sc.pos.copy()
    keys_per_step = []
    for step in range(steps):
        disp = np.full_like(pos, 0.01 * step)
        w.step(pos, sc.quat, disp, sc.shape_id, sc.flags)
        keys_per_step.append(w.pairs().copy())
        used = pos
        pos = pos + 0.03
    keys = keys_per_step[-1]
    pa = (keys >> np.uint64(32)).astype(np.uint32)
    pb = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hit, out, _ = oracle.gjk_epa_pairs(sc.shapes, used, sc.quat, sc.shape_id, pa, pb)
    return {"side": side, "steps": steps, "hit": hit, "contacts": out, **{f"keys{k}": v for k, v in enumerate(keys_per_step)}}


def pairs_mixed(n=400, seed=0x601DE5):
    sc, pa, pb = random_pairs_scene(n, seed)
    hit, out, _ = oracle.gjk_epa_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb)
    return {"n": n, "seed": seed, "hit": hit, "contacts": out}


def downstream_state(n, seed=77):
    """Seeded body state for the rows fixture (shared with tests/test_golden.py)."""
    from scenes import SplitMix64

    rng = SplitMix64(seed)
    vel = rng.uniform(-1.0, 1.0, n, 3)
    w = rng.uniform(-2.0, 2.0, n, 3)
    mass = rng.uniform(0.5, 3.0, n)
    a = rng.uniform(-0.2, 0.2, n, 3, 3)
    inertia = (np.einsum("nij,nkj->nik", a, a) + np.eye(3) * rng.uniform(0.3, 2.0, n)[:, None, None]).reshape(n, 9)
    rest = rng.uniform(0.0, 1.0, n)
    fric = rng.uniform(0.1, 1.0, n)
    origins = rng.uniform(-1.0, 7.0, 64, 3)
    dirs = rng.uniform(-1.0, 1.0, 64, 3)
    dirs[::5, 1] = 0.0
    return vel, w, mass, inertia, rest, fric, origins, dirs


def rows_c3(side=8, steps=4):
    sc = scene_c3(side=side)
    n = sc.n
    vel, w, mass, inertia, rest, fric, origins, dirs = downstream_state(n)
    world = oracle.World(sc.shapes)
    M = oracle.Manifolds()
    pos = sc.pos.copy()
    for step in range(steps):
        disp = np.full_like(pos, 0.01 * step)
        world.step(pos, sc.quat, disp, sc.shape_id, sc.flags)
        keys = world.pairs()
        pa = (keys >> np.uint64(32)).astype(np.uint32)
        pb = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        hit, out, _ = oracle.gjk_epa_pairs(sc.shapes, pos, sc.quat, sc.shape_id, pa, pb)
        M.step(keys, hit, out, pos, sc.quat)
        used = pos
        pos = pos + 0.03
    mk, mc, mp = M.get()
    rk, rp, rows = M.setup_contacts(used, sc.quat, vel, w, mass, inertia, rest, fric, 1.0 / 60.0, 9.81)
    rays = []
    for r in range(len(origins)):
        ids, d = world.raycast(origins[r], dirs[r], 6.0)
        o = np.argsort(ids, kind="stable")
        rays += [(r, int(i), float(x)) for i, x in zip(ids[o], d[o])]
    ray_arr = np.array(rays, dtype=[("ray", "<u4"), ("body", "<u4"), ("distance", "<f8")])
    dyn = oracle.Dynamics(sc.pos, sc.quat, vel, w, mass, inertia, sc.flags)
    for _ in range(3):
        dyn.integrate_velocities(1.0 / 60.0, (0.0, -9.81, 0.0))
        dyn.integrate_positions(1.0 / 60.0)
    return {"side": side, "steps": steps, "man_keys": mk, "man_counts": mc, "row_keys": rk, "row_points": rp, "rows": rows,
            "rays": ray_arr, "dyn_vel": dyn.vel, "dyn_pos": dyn.pos, "dyn_quat": dyn.quat}


if __name__ == "__main__":
    oracle.build()
    np.savez_compressed(os.path.join(HERE, "world_c3_side6.npz"), **world_c3())
    np.savez_compressed(os.path.join(HERE, "pairs_mixed.npz"), **pairs_mixed())
    np.savez_compressed(os.path.join(HERE, "rows_c3_side8.npz"), **rows_c3())
    for f in ("world_c3_side6.npz", "pairs_mixed.npz", "rows_c3_side8.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
