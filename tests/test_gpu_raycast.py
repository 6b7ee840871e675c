"""-m gpu parity of pk_raycast against the oracle's restatement of dynamic_bvh::raycast /
world_base::raycast (bvh.h:346-450, core/world.h:260-319): the set of (body, distance) entries per ray is
identical, distances bit for bit; the closest mode returns the minimum (lowest id among equals)."""
import numpy as np
import pytest

import oracle
import physkit_b200 as pk
from scenes import Scene, SplitMix64, scene_c1, scene_c2, scene_c3

pytestmark = pytest.mark.gpu


def _rays(rng, n, lo, hi):
    o = rng.uniform(lo, hi, n, 3)
    d = rng.uniform(-1, 1, n, 3)
    # a share of axis-aligned rays: zero direction components exercise the 0·inf → NaN substitution
    d[::7, 1] = 0.0
    d[::11, 2] = 0.0
    d[::13, 0] = -0.0
    return o, d


def _compare_all(hits, per_ray_ref):
    """hits: structured array sorted by (ray, body); per_ray_ref: list of (ids, dists) per ray, any order."""
    assert np.all(np.diff(hits["ray"].astype(np.int64) << 32 | hits["body"]) > 0), "not sorted by (ray, body) / duplicates"
    total = 0
    for r, (ids, d) in enumerate(per_ray_ref):
        got = hits[hits["ray"] == r]
        order = np.argsort(ids, kind="stable")
        assert np.array_equal(got["body"], ids[order]), f"ray {r}: bodies differ"
        assert np.array_equal(got["distance"].view(np.uint64), d[order].view(np.uint64)), f"ray {r}: distances differ bitwise"
        total += len(ids)
    assert total == len(hits)
    return total


def _closest_from(per_ray_ref):
    out = []
    for r, (ids, d) in enumerate(per_ray_ref):
        if len(ids) == 0:
            continue
        m = d.min()
        out.append((r, ids[d == m].min(), m))
    return out


@pytest.mark.parametrize("n", [2, 17, 5000])
def test_query_mode_rays_match_dynamic_bvh(n):
    """Exact boxes (the way tests/dynamic_bvh/main.cpp drives the tree): every ray's set equals the faithful
    dynamic_bvh::raycast over the same boxes."""
    from gpu_util import make_context

    sc = scene_c2(n, extent=12.0)
    ctx = make_context(sc, max_pairs=max(64, 200 * n))
    ctx.collide()
    boxes = oracle.bounds(sc.shapes, sc.pos, sc.quat, sc.shape_id)
    tree = oracle.DynamicBVH()
    for i in range(n):
        tree.add(i, boxes[i])
    rng = SplitMix64(77 + n)
    o, d = _rays(rng, 300, -14, 14)
    ref = [tree.raycast(o[r], d[r], 40.0) for r in range(len(o))]
    hits = ctx.raycast(o, d, 40.0)
    total = _compare_all(hits, ref)
    assert total > 0
    closest = ctx.raycast(o, d, 40.0, mode=pk.RAY_CLOSEST)
    want = _closest_from(ref)
    assert [(int(h["ray"]), int(h["body"])) for h in closest] == [(r, int(b)) for r, b, _ in want]
    assert np.array_equal(closest["distance"].view(np.uint64), np.array([m for _, _, m in want]).view(np.uint64))
    assert ctx.raycast_device_ms() > 0
    ctx.close()


def test_world_mode_rays_see_fat_boxes_static_and_dynamic():
    """world_base::raycast after a few steps: stored (fat) boxes of the dynamic tree + the static ground."""
    from gpu_util import make_context

    sc = scene_c1(side=6)
    ctx = make_context(sc, max_pairs=200_000, mode=pk.MODE_WORLD)
    w = oracle.World(sc.shapes)
    pos = sc.pos.copy()
    for step in range(4):
        disp = np.zeros_like(pos)
        disp[:, 1] = -0.02 * step
        w.step(pos, sc.quat, disp, sc.shape_id, sc.flags)
        ctx.upload(pos, sc.quat, disp, sc.shape_id, sc.flags)
        ctx.collide()
        pos = pos + np.array([0.0, -0.06, 0.0]) * ((sc.flags & 1) == 0)[:, None]
    rng = SplitMix64(4242)
    o, d = _rays(rng, 400, -6, 12)
    o[:50] = [0.1, 40.0, 0.2]
    d[:50] = rng.uniform(-0.2, 0.2, 50, 3) + np.array([0.0, -1.0, 0.0])  # lidar-style fan from above
    ref = [w.raycast(o[r], d[r], 100.0) for r in range(len(o))]
    hits = ctx.raycast(o, d, 100.0)
    total = _compare_all(hits, ref)
    assert total > 500
    ground = int(np.nonzero(sc.flags & 1)[0][0])
    assert ground in set(hits["body"].tolist())
    ctx.close()


def test_per_ray_max_distance_and_overflow_retry():
    from gpu_util import make_context

    sc = scene_c3(side=12)
    ctx = make_context(sc, max_pairs=400_000, mode=pk.MODE_WORLD)
    w = oracle.World(sc.shapes)
    w.step(sc.pos, sc.quat, sc.disp, sc.shape_id, sc.flags)
    ctx.collide()
    rng = SplitMix64(9)
    o, d = _rays(rng, 256, -1, 10)
    md = rng.uniform(0.0, 8.0, 256)
    ref = [w.raycast(o[r], d[r], md[r]) for r in range(256)]
    hits = ctx.raycast(o, d, md)
    total = _compare_all(hits, ref)
    assert total > 1000
    # a capacity that is too small reports the one to retry with and writes nothing
    with pytest.raises(pk.PkError) as e:
        ctx.raycast(o, d, md, capacity=16)
    assert e.value.status == -5
    assert len(ctx.raycast(o, d, md, capacity=total)) == total
    ctx.close()


def test_batched_worlds_rays_stay_in_their_world():
    from gpu_util import make_context

    nw = 27
    base = scene_c1(side=4, spacing=1.1)
    per = base.n
    pos = np.concatenate([base.pos + SplitMix64(500 + k).uniform(-0.05, 0.05, per, 3) for k in range(nw)])
    quat = np.tile(base.quat, (nw, 1))
    sid = np.tile(base.shape_id, nw)
    flags = np.tile(base.flags, nw)
    wid = np.repeat(np.arange(nw, dtype=np.uint32), per)
    sc = Scene(base.shapes, pos, quat, sid, flags)
    ctx = make_context(sc, max_pairs=400_000, mode=pk.MODE_WORLD, num_worlds=nw, world_id_array=wid)
    ctx.collide()
    sample = [0, 13, 26]
    worlds = {}
    for k in sample:
        worlds[k] = oracle.World(base.shapes)
        worlds[k].step(pos[k * per:(k + 1) * per], base.quat, np.zeros((per, 3)), base.shape_id, base.flags)
    rng = SplitMix64(31)
    o, d = _rays(rng, 240, -4, 8)
    rw = np.array([sample[r % 3] for r in range(240)], dtype=np.uint32)
    ref = []
    for r in range(240):
        ids, dist = worlds[int(rw[r])].raycast(o[r], d[r], 60.0)
        ref.append((ids + int(rw[r]) * per, dist))  # body ids are global in the batched context
    hits = ctx.raycast(o, d, 60.0, world=rw)
    total = _compare_all(hits, ref)
    assert total > 100
    assert np.array_equal(hits["body"] // per, rw[hits["ray"]])
    ctx.close()


def test_edge_cases_no_tree_empty_batch_and_call_order():
    ctx = pk.Context(4, 64, mode=pk.MODE_WORLD, max_shapes=2)
    sid = ctx.add_shapes([("obb", np.array([0.5, 0.5, 0.5]))])[0]
    with pytest.raises(pk.PkError) as e:  # no step yet: there is no tree
        ctx.raycast([[0, 0, 0]], [[1, 0, 0]], 10.0)
    assert e.value.status == -7
    ctx.resize(2)
    pos = np.array([[5.0, 0, 0], [9.0, 0, 0]])
    quat = np.tile([0.0, 0, 0, 1], (2, 1))
    flags = np.array([2, 0], np.uint8)  # one body alive: the step builds no tree
    ctx.upload(pos, quat, np.zeros((2, 3)), np.array([sid, sid], np.uint32), flags)
    ctx.collide()
    h = ctx.raycast([[0, 0, 0], [0, 5, 0]], [[1, 0, 0], [1, 0, 0]], 100.0)
    assert [(int(x["ray"]), int(x["body"]), float(x["distance"])) for x in h] == [(0, 0, 4.5)]
    h = ctx.raycast([[0, 0, 0]], [[1, 0, 0]], 100.0, mode=pk.RAY_CLOSEST)
    assert len(h) == 1 and h[0]["distance"] == 4.5
    assert len(ctx.raycast(np.zeros((0, 3)), np.zeros((0, 3)), 1.0)) == 0
    flags[:] = 0
    ctx.upload(pos, quat, np.zeros((2, 3)), np.array([sid, sid], np.uint32), flags)
    ctx.collide()
    assert len(ctx.raycast([[0, 0, 0]], [[1, 0, 0]], 100.0)) == 0
    ctx.close()


def test_one_million_rays_against_c3_pile():
    """Throughput-sized batch (lidar / RL sensors): 1 M rays against a 125 k-body pile; a 2 k-ray sample equals
    the oracle, and the closest hit of every ray is the minimum of its full set."""
    from gpu_util import make_context

    sc = scene_c3(side=50)
    ctx = make_context(sc, max_pairs=8_000_000, mode=pk.MODE_WORLD, max_contacts=1_000_000)
    w = oracle.World(sc.shapes)
    w.step(sc.pos, sc.quat, sc.disp, sc.shape_id, sc.flags)
    ctx.collide()
    rng = SplitMix64(123)
    n = 1_000_000
    o = rng.uniform(-2, 42, n, 3)
    d = rng.uniform(-1, 1, n, 3)
    md = np.full(n, 1.0)
    hits = ctx.raycast(o, d, md)
    ms_all = ctx.raycast_device_ms()
    closest = ctx.raycast(o, d, md, mode=pk.RAY_CLOSEST)
    ms_closest = ctx.raycast_device_ms()
    print(f"1M rays: {len(hits)} entries in {ms_all:.2f} ms, closest {len(closest)} in {ms_closest:.2f} ms")
    s = 2000
    ref = [w.raycast(o[r], d[r], 1.0) for r in range(s)]
    _compare_all(hits[hits["ray"] < s], ref)
    # closest == per-ray minimum of the full set, lowest id among equals
    order = np.lexsort((hits["body"], hits["distance"], hits["ray"]))
    hs = hits[order]
    first = np.ones(len(hs), bool)
    first[1:] = hs["ray"][1:] != hs["ray"][:-1]
    want = hs[first]
    assert np.array_equal(closest["ray"], want["ray"])
    assert np.array_equal(closest["body"], want["body"])
    assert np.array_equal(closest["distance"].view(np.uint64), want["distance"].view(np.uint64))
    ctx.close()
