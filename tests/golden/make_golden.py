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


if __name__ == "__main__":
    oracle.build()
    np.savez_compressed(os.path.join(HERE, "world_c3_side6.npz"), **world_c3())
    np.savez_compressed(os.path.join(HERE, "pairs_mixed.npz"), **pairs_mixed())
    for f in ("world_c3_side6.npz", "pairs_mixed.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
