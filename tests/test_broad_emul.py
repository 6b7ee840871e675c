"""The broadphase kernels (physkit_b200/csrc/pk_broadphase.cuh: bounds + fat rule, Morton keys, LBVH hierarchy with
fused refit, ropes, self-overlap traversal, pair rows) run on the host through tests/emul.py in the order
pk_collide_resident launches them, against the oracle's faithful incremental dynamic_bvh (reference
collision_phases.h:330-445, src/bvh.cpp:239-514): stored boxes bit for bit, moved counts and the sorted pair set, every
step.  The same comparisons run through the C ABI on the GPU (tests/test_gpu_broadphase.py); this file makes them
available in the container that has none."""
import numpy as np
import pytest

import emul
import oracle
from scenes import Scene, SplitMix64, scene_c1, scene_c2, scene_c3

pytestmark = pytest.mark.skipif(not emul.available(), reason="CUDA headers not installed")


def _replay(sc, steps, mutate, rows=True):
    w = oracle.World(sc.shapes)
    g = emul.BroadWorld(sc.shapes, sc.n, rows=rows)
    pos, quat, flags = sc.pos.copy(), sc.quat.copy(), sc.flags.copy()
    total = 0
    for step in range(steps):
        disp = mutate(step, pos, quat, flags)
        moved = w.step(pos, quat, disp, sc.shape_id, flags)
        assert g.step(pos, quat, disp, sc.shape_id, flags) == moved, f"step {step}"
        want = w.pairs()
        assert np.array_equal(g.pairs(), want), f"step {step}: {g.npairs} vs {len(want)}"
        got = g.stored()
        for i in np.nonzero(flags & 2)[0][:: max(1, sc.n // 64)]:
            assert np.array_equal(got[i].view(np.uint64), w.stored(int(i)).view(np.uint64)), f"step {step} body {i}"
        total += len(want)
        dyn = ((flags & 1) == 0) & ((flags & 2) != 0)
        pos += disp * dyn[:, None]
    return total


def test_world_replay_c1_style_with_bodies_created_and_destroyed():
    """Ground + lattice of box hulls falling under gravity; the first step yields no pair (collision_phases.h:342-346),
    bodies appear at step 5, disappear at 12 and come back at 20 (arena slots are reused, core/world.h:205-206)."""
    sc = scene_c1(side=4, spacing=1.05)
    n = sc.n
    rng = SplitMix64(3)
    vel = np.zeros((n, 3))
    late = np.zeros(n, dtype=bool)
    late[5::9] = True
    sc.flags[late] = 0
    dt = 1.0 / 60.0

    def mutate(step, pos, quat, flags):
        if step == 5:
            flags[late] = 2
        if step == 12:
            flags[7:60:6] = 0
        if step == 20:
            flags[7:60:6] = 2
        dyn = (flags & 1) == 0
        vel[dyn, 1] -= 9.81 * dt
        if step % 4 == 0:
            vel[dyn] += rng.uniform(-0.3, 0.3, n, 3)[dyn]
        return vel * dt

    assert _replay(sc, 24, mutate) > 150


@pytest.mark.parametrize("rows,side", [(True, 9), (False, 7)])
def test_world_replay_c3_style_rows_and_list_forms(rows, side):
    """Spheres and boxes drifting at random: pairs collected in per-body rows and written out sorted, or appended to
    a list and radix-sorted (several tiles: one radix_hist / radix_scan / radix_scatter launch per pass) — the same set
    either way."""
    sc = scene_c3(side=side)
    rng = SplitMix64(9)
    assert _replay(sc, 6, lambda step, pos, quat, flags: rng.uniform(-0.08, 0.08, sc.n, 3), rows=rows) > (10_000 if side == 9 else 5_000)


@pytest.mark.parametrize("n", [2, 3, 17, 1000, 6000])
def test_query_mode_is_the_static_pose_pair_set(n):
    """BASELINE C2 shape: exact boxes, every overlapping pair (dynamic_bvh add all + query all, and brute force)."""
    sc = scene_c2(n, extent=50.0 * (max(n, 64) / 100_000.0) ** (1 / 3) * 0.6)
    g = emul.BroadWorld(sc.shapes, n, mode_query=True)
    g.step(sc.pos, sc.quat, sc.disp, sc.shape_id, sc.flags)
    boxes = oracle.bounds(sc.shapes, sc.pos, sc.quat, sc.shape_id)
    assert np.array_equal(g.stored().view(np.uint64), boxes.view(np.uint64))
    assert np.array_equal(g.pairs(), oracle.query_pairs(boxes))
    if n <= 1000:
        assert np.array_equal(g.pairs(), oracle.brute_pairs(boxes))


def test_degenerate_boxes_touching_corners_and_coincident_bodies():
    """aabb::intersects is inclusive (bounds.h:87-92; tests/mesh/main.cpp:109-121): corner-touching, flat and coincident
    boxes pair; equal Morton keys must still give a proper tree."""
    shapes = [("aabb", (0, 0, 0), (1, 1, 1)), ("aabb", (1, 1, 1), (2, 2, 2)), ("aabb", (0, 0, 0), (1, 1, 1)), ("aabb", (0.5, 0.5, 1), (3, 3, 1)),
              ("aabb", (5, 5, 5), (6, 6, 6))] + [("aabb", (10, 10, 10), (11, 11, 11))] * 40
    n = len(shapes)
    sc = Scene(shapes, np.zeros((n, 3)), np.tile([0, 0, 0, 1.0], (n, 1)), np.arange(n))
    g = emul.BroadWorld(sc.shapes, n, mode_query=True)
    g.step(sc.pos, sc.quat, sc.disp, sc.shape_id, sc.flags)
    boxes = oracle.bounds(sc.shapes, sc.pos, sc.quat, sc.shape_id)
    want = oracle.brute_pairs(boxes)
    assert np.array_equal(g.pairs(), want) and len(want) == 6 + 40 * 39 // 2


def test_sharded_traversal_partitions_the_pair_set():
    """One world over N ranks: rank r traverses the sorted leaves [m·r/N, m·(r+1)/N) and a leaf only reports partners to
    its right, so the ranks' sets are disjoint and their union is the whole set."""
    sc = scene_c3(side=8)
    whole = emul.BroadWorld(sc.shapes, sc.n, mode_query=True)
    whole.step(sc.pos, sc.quat, sc.disp, sc.shape_id, sc.flags)
    parts = []
    for r in range(3):
        g = emul.BroadWorld(sc.shapes, sc.n, mode_query=True, shard=(r, 3))
        g.step(sc.pos, sc.quat, sc.disp, sc.shape_id, sc.flags)
        parts.append(g.pairs())
        assert np.all(np.diff(parts[-1].astype(np.int64)) > 0)
    allp = np.concatenate(parts)
    assert len(np.unique(allp)) == len(allp) and np.array_equal(np.sort(allp), whole.pairs())
    assert min(len(p) for p in parts) > 0.15 * len(allp)


def test_batched_worlds_do_not_interact():
    """BASELINE C5 shape: independent worlds in one context; pairs form inside a world only and equal the per-world oracle."""
    nw = 5
    base = scene_c1(side=4, spacing=0.95)
    per = base.n
    pos = np.concatenate([base.pos + SplitMix64(100 + k).uniform(-0.02, 0.02, per, 3) for k in range(nw)])
    quat, sid, flags = np.tile(base.quat, (nw, 1)), np.tile(base.shape_id, nw), np.tile(base.flags, nw)
    wid = np.repeat(np.arange(nw, dtype=np.uint32), per)
    g = emul.BroadWorld(base.shapes, nw * per, num_worlds=nw)
    ws = [oracle.World(base.shapes) for _ in range(nw)]
    for step in range(3):
        p = pos + np.array([0.0, -0.03 * step, 0.0]) * (flags[:, None] == 2)
        disp = np.zeros_like(p)
        g.step(p, quat, disp, sid, flags, wid)
        want = []
        for k, w in enumerate(ws):
            sl = slice(k * per, (k + 1) * per)
            w.step(p[sl], quat[sl], disp[sl], base.shape_id, base.flags)
            keys = w.pairs()
            off = np.uint64(k * per)
            want.append(((keys >> np.uint64(32)) + off) << np.uint64(32) | ((keys & np.uint64(0xFFFFFFFF)) + off))
        assert np.array_equal(g.pairs(), np.sort(np.concatenate(want))), f"step {step}"
    assert g.npairs > 300


def test_a_full_pair_row_is_reported():
    """More than 64 partners with a larger id: the row overflows and the step has to be repeated in list form (pk_api.cu
    repeat_step); the list form gives the set."""
    n = 80
    shapes = [("aabb", (0, 0, 0), (1, 1, 1))] * n
    sc = Scene(shapes, np.zeros((n, 3)), np.tile([0, 0, 0, 1.0], (n, 1)), np.arange(n))
    g = emul.BroadWorld(sc.shapes, n, mode_query=True, rows=True)
    with pytest.raises(RuntimeError, match="row overflowed"):
        g.step(sc.pos, sc.quat, sc.disp, sc.shape_id, sc.flags)
    g = emul.BroadWorld(sc.shapes, n, mode_query=True, rows=False)
    g.step(sc.pos, sc.quat, sc.disp, sc.shape_id, sc.flags)
    assert g.npairs == n * (n - 1) // 2 and np.all(np.diff(g.pairs().astype(np.int64)) > 0)
