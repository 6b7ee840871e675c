"""GPU parity: LBVH broadphase + full collision stage vs the CPU oracle, through the C ABI.

Bar: the candidate pair set is bit-exact as a sorted (i,j) list; stored (fat) boxes are
bit-identical; contacts are bit-identical."""
import numpy as np
import pytest

import oracle
import physkit_b200 as pk
from scenes import Scene, SplitMix64, scene_c1, scene_c2, scene_c3

pytestmark = pytest.mark.gpu


def _split(keys):
    return (keys >> np.uint64(32)).astype(np.uint32), (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def _check_contacts(sc, pos, quat, keys, contacts):
    from gpu_util import contacts_equal_bitwise

    pa, pb = _split(keys)
    hit_ref, out_ref, _ = oracle.gjk_epa_pairs(sc.shapes, pos, quat, sc.shape_id, pa, pb, nthreads=8)
    m = hit_ref.astype(bool)
    assert np.array_equal(contacts["key"], keys[m])
    contacts_equal_bitwise(contacts, np.ones(len(contacts), np.uint8), np.ones(len(contacts), np.uint8), out_ref[m])


@pytest.mark.parametrize("n", [2, 3, 17, 1000, 20_000])
def test_query_mode_pairs_exact(n):
    """BASELINE C2 shape (random OBBs), static-pose mode: pair set == faithful dynamic_bvh + == brute force."""
    from gpu_util import make_context

    sc = scene_c2(n, extent=50.0 * (max(n, 64) / 100_000.0) ** (1 / 3) * 0.6)
    ctx = make_context(sc, max_pairs=max(64, 40 * n))
    res = ctx.collide()
    keys = ctx.pairs()
    boxes = oracle.bounds(sc.shapes, sc.pos, sc.quat, sc.shape_id)
    assert np.array_equal(ctx.stored_bounds(0, n).view(np.uint64), boxes.view(np.uint64))
    want = oracle.query_pairs(boxes)
    assert res.num_pairs == len(want)
    assert np.array_equal(keys, want)
    if n <= 1000:
        assert np.array_equal(keys, oracle.brute_pairs(boxes))
    _check_contacts(sc, sc.pos, sc.quat, keys, ctx.contacts())
    ctx.close()


def test_query_mode_c2_full_size():
    """BASELINE config C2 at full size: 100 k randomly posed OBBs in a 100 m cube, exact match."""
    from gpu_util import make_context

    sc = scene_c2(100_000)
    ctx = make_context(sc, max_pairs=2_000_000)
    res = ctx.collide()
    keys = ctx.pairs()
    boxes = oracle.bounds(sc.shapes, sc.pos, sc.quat, sc.shape_id)
    want = oracle.query_pairs(boxes)
    assert np.array_equal(keys, want)
    assert res.num_pairs > 100_000
    _check_contacts(sc, sc.pos, sc.quat, keys, ctx.contacts())
    ctx.close()


def test_degenerate_inputs():
    """Empty scene, one body, coincident boxes, zero-volume boxes, corner touch (inclusive)."""
    ctx = pk.Context(64, 4096, mode=pk.MODE_QUERY, max_shapes=64)
    ids = ctx.add_shapes([("obb", (0.5, 0.5, 0.5)), ("obb", (0.5, 0.5, 0.0))])
    ctx.resize(0)
    assert ctx.collide().num_pairs == 0 and len(ctx.pairs()) == 0
    I = [0.0, 0.0, 0.0, 1.0]
    ctx.resize(1)
    ctx.upload([[0, 0, 0]], [I], None, [ids[0]], [2])
    assert ctx.collide().num_pairs == 0
    # 5 coincident boxes → all 10 pairs (tests/dynamic_bvh/main.cpp:704-718)
    ctx.resize(5)
    ctx.upload(np.zeros((5, 3)), [I] * 5, None, [ids[0]] * 5, [2] * 5)
    assert ctx.collide().num_pairs == 10
    # corner touch counts (tests/mesh/main.cpp:118-120); flat box (main.cpp:687-702)
    ctx.resize(3)
    ctx.upload([[0, 0, 0], [1, 1, 1], [0.2, 0.2, 0.5]], [I] * 3, None, [ids[0], ids[0], ids[1]], [2] * 3)
    ctx.collide()
    assert list(ctx.pairs()) == [(0 << 32) | 1, (0 << 32) | 2, (1 << 32) | 2]
    # dead bodies never pair
    ctx.upload([[0, 0, 0], [1, 1, 1], [0.2, 0.2, 0.5]], [I] * 3, None, [ids[0], ids[0], ids[1]], [2, 0, 2])
    ctx.collide()
    assert list(ctx.pairs()) == [(0 << 32) | 2]
    ctx.close()


def test_pair_overflow_is_reported():
    ctx = pk.Context(64, 4, mode=pk.MODE_QUERY, max_shapes=4)
    sid = ctx.add_shape(("obb", (0.5, 0.5, 0.5)))
    ctx.resize(8)
    ctx.upload(np.zeros((8, 3)), [[0, 0, 0, 1.0]] * 8, None, [sid] * 8, [2] * 8)
    with pytest.raises(pk.PkError) as e:
        ctx.collide()
    assert e.value.status == -5
    assert ctx.result.pairs_required == 28
    ctx.close()


def _world_replay(sc, steps, mutate, max_pairs, **ctxkw):
    """Drive oracle.World (faithful incremental broadphase) and the GPU ctx with the same inputs."""
    from gpu_util import make_context

    ctx = make_context(sc, max_pairs=max_pairs, mode=pk.MODE_WORLD, **ctxkw)
    w = oracle.World(sc.shapes)
    pos, quat, flags = sc.pos.copy(), sc.quat.copy(), sc.flags.copy()
    total = 0
    for step in range(steps):
        disp = mutate(step, pos, quat, flags)
        moved = w.step(pos, quat, disp, sc.shape_id, flags)
        ctx.upload(pos, quat, disp, sc.shape_id, flags)
        res = ctx.collide()
        keys = ctx.pairs()
        want = w.pairs()
        assert res.num_moved == moved, f"step {step}"
        assert np.array_equal(keys, want), f"step {step}: {len(keys)} vs {len(want)}"
        alive = np.nonzero(flags & 2)[0]
        got_boxes = ctx.stored_bounds(0, sc.n)
        for i in alive[:: max(1, len(alive) // 64)]:
            assert np.array_equal(got_boxes[i].view(np.uint64), w.stored(int(i)).view(np.uint64))
        if step % 7 == 3 and len(keys):
            _check_contacts(sc, pos, quat, keys, ctx.contacts())
        total += len(keys)
        dyn = ((flags & 1) == 0) & ((flags & 2) != 0)
        pos += disp * dyn[:, None]
    ctx.close()
    return total


def test_world_mode_replay_c1_style():
    """BASELINE C1 shape: ground + lattice of 8-vertex box hulls falling under gravity (no solver:
    contact response is out of scope, bodies interpenetrate), 60 steps at 60 Hz.  Step 1 must yield
    zero pairs (first-step quirk), bodies are created and destroyed mid-run."""
    sc = scene_c1(side=6, spacing=1.05)
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

    total = _world_replay(sc, 60, mutate, max_pairs=200_000)
    assert total > 1000


def test_world_mode_first_step_has_no_pairs():
    from gpu_util import make_context

    sc = scene_c3(side=8)
    ctx = make_context(sc, max_pairs=100_000, mode=pk.MODE_WORLD)
    assert ctx.collide().num_pairs == 0  # collision_phases.h:342-346 + src/world.cpp:30
    ctx.update_pose(sc.pos + 0.01)
    res = ctx.collide()
    assert res.num_moved == sc.n and res.num_pairs > 0
    ctx.close()


def test_world_mode_c3_style_mixed_shapes():
    """C3 shapes (analytic spheres + OBBs) in world mode with random drift."""
    sc = scene_c3(side=12)
    rng = SplitMix64(9)

    def mutate(step, pos, quat, flags):
        return rng.uniform(-0.08, 0.08, sc.n, 3)

    total = _world_replay(sc, 12, mutate, max_pairs=400_000)
    assert total > 10_000


def test_batched_worlds_do_not_interact():
    """BASELINE C5 shape: independent worlds in one ctx; pairs only form inside a world and equal
    the per-world oracle result (ids offset)."""
    from gpu_util import make_context

    nw = 9
    base = scene_c1(side=4, spacing=0.95)
    per = base.n
    pos = np.concatenate([base.pos + SplitMix64(100 + k).uniform(-0.02, 0.02, per, 3) for k in range(nw)])
    quat = np.tile(base.quat, (nw, 1))
    sid = np.tile(base.shape_id, nw)
    flags = np.tile(base.flags, nw)
    wid = np.repeat(np.arange(nw, dtype=np.uint32), per)
    sc = Scene(base.shapes, pos, quat, sid, flags)
    ctx = make_context(sc, max_pairs=400_000, mode=pk.MODE_WORLD, num_worlds=nw, world_id_array=wid)
    worlds = [oracle.World(base.shapes) for _ in range(nw)]
    p = pos.copy()
    for step in range(4):
        disp = np.full((len(p), 3), 0.0)
        disp[:, 1] = -0.05 * step
        ctx.upload(p, quat, disp, sid, flags, wid)
        ctx.collide()
        keys = ctx.pairs()
        want = []
        for k, w in enumerate(worlds):
            sl = slice(k * per, (k + 1) * per)
            w.step(p[sl], quat[sl], disp[sl], base.shape_id, base.flags)
            kk = w.pairs()
            a, b = _split(kk)
            want.append(((a.astype(np.uint64) + np.uint64(k * per)) << np.uint64(32)) | (b.astype(np.uint64) + np.uint64(k * per)))
        want = np.sort(np.concatenate(want))
        assert np.array_equal(keys, want), f"step {step}"
        p = p + disp * ((flags & 1) == 0)[:, None]
    assert len(keys) > 0
    ctx.close()


@pytest.mark.parametrize("shards", [2, 3, 8])
def test_pair_sharding_partitions_the_set(shards):
    """One large world, pairs partitioned across ranks by sorted-leaf range: the shards are disjoint
    and their union is the full set; contacts likewise."""
    from gpu_util import make_context

    sc = scene_c3(side=14)
    full = make_context(sc, max_pairs=600_000)
    full.collide()
    keys_full, con_full = full.pairs(), full.contacts()
    full.close()
    parts, cparts = [], []
    for r in range(shards):
        ctx = make_context(sc, max_pairs=600_000, shard_rank=r, shard_count=shards)
        ctx.collide()
        parts.append(ctx.pairs())
        cparts.append(ctx.contacts())
        ctx.close()
    allk = np.concatenate(parts)
    assert len(allk) == len(keys_full) and len(np.unique(allk)) == len(allk)
    assert np.array_equal(np.sort(allk), keys_full)
    allc = np.concatenate(cparts)
    allc = allc[np.argsort(allc["key"], kind="stable")]
    assert np.array_equal(allc.view(np.uint8), con_full.view(np.uint8))
    assert min(len(p) for p in parts) > 0.3 * len(keys_full) / shards  # no empty / degenerate shard


def test_contact_points_in_body_frames():
    """narrow_phase::calculate stores contact_point(info, a, b) (collision_phases.h:75-88, 257-263): the witness
    points in the bodies' own frames.  pk_contact_points against the oracle's restatement, bit for bit."""
    from gpu_util import make_context

    sc = scene_c3(side=12)
    ctx = make_context(sc, max_pairs=400_000, mode=pk.MODE_QUERY)
    try:
        ctx.collide()
        con = ctx.contacts()
        pts = ctx.contact_points()
    finally:
        ctx.close()
    assert len(con) > 500 and pts.shape == (len(con), 6)
    pa, pb = _split(con["key"])
    c10 = np.concatenate([con["normal"], con["world_a"], con["world_b"], con["depth"][:, None]], axis=1)
    want = oracle.contact_points(sc.pos, sc.quat, pa, pb, c10)
    assert np.array_equal(pts.view(np.uint64), want.view(np.uint64))
    # and they are what they claim to be: rotating back and translating returns the world points
    from scipy.spatial.transform import Rotation

    back = Rotation.from_quat(sc.quat[pa]).apply(pts[:, :3]) + sc.pos[pa]
    assert np.allclose(back, con["world_a"], atol=1e-12)


def test_pk_create_multi_shards_one_world_over_contexts():
    """pk_create_multi / pk_multi_collide: n contexts from one process, context i traversing its
    slice of the sorted leaves.  Run with the same device listed three times: the shards' pair and contact sets are
    disjoint and their union is what one context produces, bit for bit."""
    from gpu_util import make_context

    sc = scene_c3(side=20)
    w = oracle.World(sc.shapes)
    one = make_context(sc, max_pairs=400_000, mode=pk.MODE_WORLD)
    m = pk.MultiContext([0, 0, 0], sc.n, 400_000, mode=pk.MODE_WORLD, max_shapes=len(sc.shapes))
    m.add_shapes(sc.shapes)
    m.resize(sc.n)
    pos = sc.pos.copy()
    for step in range(3):
        disp = np.full_like(pos, 0.01 * step)
        w.step(pos, sc.quat, disp, sc.shape_id, sc.flags)
        one.upload(pos, sc.quat, disp, sc.shape_id, sc.flags)
        one.collide()
        m.upload(pos, sc.quat, disp, sc.shape_id, sc.flags)
        tot = m.collide()
        keys = [c.pairs() for c in m.ctx]
        cons = [c.contacts() for c in m.ctx]
        allk = np.concatenate(keys)
        assert tot.num_pairs == len(allk) == len(np.unique(allk))
        assert np.array_equal(np.sort(allk), w.pairs())
        assert np.array_equal(np.sort(allk), one.pairs())
        allc = np.concatenate(cons)
        assert tot.num_contacts == len(allc)
        order = np.argsort(allc["key"], kind="stable")
        assert np.array_equal(allc[order].view(np.uint8), one.contacts().view(np.uint8))
        pos = pos + 0.03
    assert tot.num_pairs > 10_000 and min(len(k) for k in keys) > 0
    one.close()
    m.close()


def test_pair_overflow_is_recoverable():
    """PK_E_PAIR_OVERFLOW (pairs, then GJK hits): pk_reserve_pairs and the SAME step again give what a context with enough
    room gives — pair set, contacts, stored boxes — and later steps stay in step with the oracle (the failed attempts
    did not advance the epoch).  Also the ADVICE r1 case: more GJK hits than contact records must not write past them."""
    from gpu_util import make_context

    sc = scene_c3(side=12)
    big = make_context(sc, max_pairs=200_000, mode=pk.MODE_WORLD)
    cap_pairs, cap_contacts = 3_000, 100
    small = make_context(sc, max_pairs=cap_pairs, mode=pk.MODE_WORLD, max_contacts=cap_contacts)
    w = oracle.World(sc.shapes)
    pos = sc.pos.copy()
    total_tries = 0
    for step in range(4):
        disp = np.full_like(pos, 0.01 * step)
        w.step(pos, sc.quat, disp, sc.shape_id, sc.flags)
        big.upload(pos, sc.quat, disp, sc.shape_id, sc.flags)
        small.upload(pos, sc.quat, disp, sc.shape_id, sc.flags)
        big.collide()
        for attempt in range(4):
            try:
                small.collide()
                break
            except pk.PkError as e:
                assert e.status == pk.PK_E_PAIR_OVERFLOW and attempt < 3
                total_tries += 1
                need = int(small.result.pairs_required)
                if need:  # candidate pairs did not fit
                    cap_pairs = need + 16
                else:  # GJK hits did not fit the contact records
                    cap_contacts = int(small.result.gjk_hits) + 16
                small.reserve_pairs(cap_pairs, cap_contacts)
        assert np.array_equal(small.pairs(), w.pairs())
        assert np.array_equal(small.pairs(), big.pairs())
        assert np.array_equal(small.contacts().view(np.uint8), big.contacts().view(np.uint8))
        assert np.array_equal(small.stored_bounds(0, sc.n).view(np.uint64), big.stored_bounds(0, sc.n).view(np.uint64))
        pos = pos + 0.03
    assert total_tries >= 2  # the pairs overflowed, then the contacts
    big.close()
    small.close()


def test_resident_steps_without_the_mid_step_read_back():
    """pk_collide_resident sizes the pair sort and the narrowphase from the previous step's pair count and reads the
    count on the device.  Pair counts that grow slowly, jump by far more than the 1/8 + 64 Ki of slack (the step is run
    again with the count read back), shrink and drop to nothing must give what pk_collide gives, step after step."""
    from gpu_util import make_context

    sc = scene_c3(side=44)  # 85 k bodies, ~1.2 M pairs at spacing 0.8
    ref = make_context(sc, max_pairs=12_000_000, mode=pk.MODE_WORLD)
    res = make_context(sc, max_pairs=12_000_000, mode=pk.MODE_WORLD)
    centre = sc.pos.mean(axis=0)
    counts = []
    for step, scale in enumerate([1.0, 1.0, 0.99, 0.98, 0.62, 0.63, 1.0, 1.0, 40.0, 1.0]):
        pos = centre + (sc.pos - centre) * scale + 0.01 * step
        disp = np.zeros_like(pos)
        for c in (ref, res):
            c.upload(pos, sc.quat, disp, sc.shape_id, sc.flags)
        r0 = ref.collide(allow_epa_overflow=True)
        r1 = res.collide_resident(allow_epa_overflow=True)
        counts.append(int(r1.num_pairs))
        assert (r1.num_pairs, r1.num_contacts, r1.num_moved, r1.step_index) == (r0.num_pairs, r0.num_contacts, r0.num_moved, r0.step_index)
        res.fetch()
        assert np.array_equal(res.pairs(), ref.pairs())
        assert np.array_equal(res.contacts().view(np.uint8), ref.contacts().view(np.uint8))
    assert counts[0] == 0 and counts[1] > 1_000_000  # (the first step of a world reports no pairs)
    assert counts[4] > 2 * counts[3] + 65536 and counts[8] == 0 and counts[9] > 1_000_000, counts
    ref.close()
    res.close()


def test_pair_rows_overflow_falls_back_to_the_sorted_list(monkeypatch):
    """The pairs of a step are collected in per-body rows of 64 partners (pk_broadphase.cuh, ROWS) and written out row by
    row; a body with more partners of a larger id than a row holds — a ground slab under 300 bodies, with the lowest
    id — makes the step run again with the pair list sorted by the radix sort.  Same pairs and contacts as the oracle
    and as a context that never uses rows (PK_PAIR_RADIX=1), on the step that overflows, on the ones after it, through
    pk_collide and through pk_collide_resident."""
    from gpu_util import make_context

    rng = SplitMix64(77)
    nb = 300
    shapes = [("obb", [40.0, 40.0, 0.5])] + [("sphere", 0.4), ("obb", [0.3, 0.4, 0.5])]
    pos = np.zeros((nb + 1, 3))
    pos[1:, :2] = rng.uniform(-30.0, 30.0, nb, 2)
    pos[1:, 2] = 0.8
    quat = rng.quats(nb + 1)
    quat[0] = [0, 0, 0, 1]
    sid = np.array([0] + [1 + (k % 2) for k in range(nb)], dtype=np.uint32)
    sc = Scene(shapes, pos, quat, sid)
    rows = make_context(sc, max_pairs=100_000, mode=pk.MODE_WORLD)
    res = make_context(sc, max_pairs=100_000, mode=pk.MODE_WORLD)
    monkeypatch.setenv("PK_PAIR_RADIX", "1")
    radix = make_context(sc, max_pairs=100_000, mode=pk.MODE_WORLD)
    monkeypatch.delenv("PK_PAIR_RADIX")
    w = oracle.World(sc.shapes)
    for step in range(4):
        p = sc.pos + 0.02 * step
        disp = np.zeros_like(p)
        w.step(p, sc.quat, disp, sc.shape_id, sc.flags)
        for c in (rows, res, radix):
            c.upload(p, sc.quat, disp, sc.shape_id, sc.flags)
        r0, r1 = radix.collide(), rows.collide()
        r2 = res.collide_resident()
        res.fetch()
        assert (r0.num_pairs, r0.num_contacts, r0.num_moved) == (r1.num_pairs, r1.num_contacts, r1.num_moved) == (r2.num_pairs, r2.num_contacts, r2.num_moved)
        assert np.array_equal(rows.pairs(), w.pairs())
        if step:
            assert r0.num_pairs > nb  # the slab against everything, and some neighbours
        for c in (rows, res):
            assert np.array_equal(c.pairs(), radix.pairs())
            assert np.array_equal(c.contacts().view(np.uint8), radix.contacts().view(np.uint8))
    for c in (rows, res, radix):
        c.close()
