"""Contact manifolds: narrow_phase::calculate's merge of each pair's new contact into the
manifold kept from the previous step (reference collision_phases.h:90-320).

CPU: properties of the oracle's restatement (oracle/pk_oracle.hpp: manifold_t, manifold_merge) that the
reference's code implies.  GPU: pk_manifolds_update against that restatement over a multi-step world replay,
bit for bit, including the cached impulses a constraint solver would write back and the began / ended events."""
import numpy as np
import pytest

import oracle
from scenes import scene_c3


def _replay(steps, side, step_fn):
    """Moving pile: bodies drift a little every step so that contacts persist, slide, break and re-form."""
    sc = scene_c3(side=side)
    pos = sc.pos.copy()
    drift = (np.array([0.012, -0.007, 0.009]) * (((np.arange(sc.n) * 2654435761) >> 7) % 5 - 2)[:, None]).astype(np.float64)
    for step in range(steps):
        disp = np.zeros_like(pos)
        step_fn(step, sc, pos, disp)
        pos = pos + drift * (1.0 if step % 7 != 6 else -3.0)


def _fake_solver_impulses(keys, counts, step):
    """Deterministic stand-in for the constraint solver's accumulated impulses (constraint.h:1107-1201)."""
    m = len(keys)
    j = np.arange(4)[None, :, None]
    c = np.arange(3)[None, None, :]
    k = (keys % np.uint64(1009)).astype(np.float64)[:, None, None]
    return 0.001 * (k + 1.0) * (j + 1.0) + 0.01 * c + 0.1 * step + np.zeros((m, 4, 3))


def _oracle_step(w, M, sc, pos, disp):
    w.step(pos, sc.quat, disp, sc.shape_id, sc.flags)
    keys = w.pairs()
    pa = (keys >> np.uint64(32)).astype(np.uint32)
    pb = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hit, out, _ = oracle.gjk_epa_pairs(sc.shapes, pos, sc.quat, sc.shape_id, pa, pb, nthreads=8)
    began, ended = M.step(keys, hit, out, pos, sc.quat)
    return keys, hit, began, ended


def test_oracle_manifolds_accumulate_reduce_and_break():
    seen = {"four": 0, "ended": 0, "began": 0, "steps": 0}
    w = None
    M = oracle.Manifolds()

    def step(k, sc, pos, disp):
        nonlocal w
        if w is None:
            w = oracle.World(sc.shapes)
        keys, hit, began, ended = _oracle_step(w, M, sc, pos, disp)
        mk, mc, mp = M.get()
        assert np.all(np.diff(mk.astype(np.int64)) > 0)          # one manifold per key, sorted
        assert np.all((mc >= 1) & (mc <= 4))                      # manifold::max_contact_points
        assert np.all(np.isin(mk, keys))                          # a manifold needs its pair in the pair set
        assert np.all(np.isin(keys[hit.astype(bool)], mk))        # a new contact is always added (:312)
        # every kept point still satisfies the breaking threshold it was tested with (:300)
        for j in range(4):
            used = mc > j
            assert np.all(mp[used, j, 9] > -0.05)
        seen["four"] += int((mc == 4).sum())
        seen["began"] += len(began)
        seen["ended"] += len(ended)
        seen["steps"] += 1
        M.set_impulses(_fake_solver_impulses(mk, mc, k))

    _replay(12, 8, step)
    assert seen["four"] > 50 and seen["began"] > 300 and seen["ended"] > 5


def test_oracle_add_reduce_keeps_the_deepest_point():
    """manifold::add_reduce (:139-198): the deepest of the five candidates always survives, in first position."""
    lib = oracle.lib()
    M = oracle.Manifolds()
    key = np.array([(1 << 32) | 2], dtype=np.uint64)
    pos = np.zeros((3, 3))
    quat = np.tile([0.0, 0.0, 0.0, 1.0], (3, 1))
    depths = [0.010, 0.030, 0.020, 0.015, 0.040]
    for t, d in enumerate(depths):
        # contacts far apart on body a (so none warm-starts another), all inside the drift / breaking thresholds
        wa = np.array([0.1 * t, 0.02 * t * t, 0.0])
        c = np.concatenate([[0.0, 0.0, 1.0], wa, wa + [0.0, 0.0, d], [d]])[None, :]
        M.step(key, np.array([1], np.uint8), c, pos, quat)
    mk, mc, mp = M.get()
    assert mc[0] == 4
    assert mp[0, 0, 9] == max(depths)


@pytest.mark.gpu
def test_gpu_manifolds_match_the_oracle_over_a_replay():
    import physkit_b200 as pk
    from gpu_util import make_context

    state = {}
    M = oracle.Manifolds()
    totals = {"four": 0, "began": 0, "ended": 0}

    def step(k, sc, pos, disp):
        if "ctx" not in state:
            state["w"] = oracle.World(sc.shapes)
            state["ctx"] = make_context(sc, max_pairs=200_000, mode=pk.MODE_WORLD)
            state["ctx"].manifolds_enable(20_000)
        ctx = state["ctx"]
        keys, hit, began, ended = _oracle_step(state["w"], M, sc, pos, disp)
        ctx.upload(pos, sc.quat, disp, sc.shape_id, sc.flags)
        ctx.collide()
        assert np.array_equal(ctx.pairs(), keys)
        n_man, n_beg, n_end, _ = ctx.manifolds_update()
        got = ctx.manifolds()
        gb, ge = ctx.manifold_events()
        mk, mc, mp = M.get()
        assert n_man == len(mk) == len(got) and n_beg == len(began) and n_end == len(ended)
        assert np.array_equal(got["key"], mk) and np.array_equal(got["count"], mc)
        assert np.array_equal(gb, began) and np.array_equal(ge, ended)
        flat = np.concatenate([got["points"]["normal"], got["points"]["local_a"], got["points"]["local_b"],
                               got["points"]["depth"][..., None], got["points"]["normal_impulse"][..., None],
                               got["points"]["tangent_impulses"]], axis=2)
        assert np.array_equal(np.ascontiguousarray(flat).view(np.uint64), np.ascontiguousarray(mp).view(np.uint64)), f"step {k}"
        imp = _fake_solver_impulses(mk, mc, k)
        M.set_impulses(imp)
        ctx.manifolds_set_impulses(imp)
        totals["four"] += int((mc == 4).sum())
        totals["began"] += len(began)
        totals["ended"] += len(ended)

    try:
        _replay(14, 9, step)
    finally:
        if "ctx" in state:
            state["ctx"].close()
    assert totals["four"] > 50 and totals["began"] > 300 and totals["ended"] > 5


@pytest.mark.gpu
def test_gpu_manifold_state_errors():
    import physkit_b200 as pk
    from gpu_util import make_context

    sc = scene_c3(side=4)
    ctx = make_context(sc, max_pairs=10_000, mode=pk.MODE_WORLD)
    try:
        with pytest.raises(pk.PkError):
            ctx.manifolds_update()  # not enabled
        ctx.manifolds_enable(1000)
        with pytest.raises(pk.PkError):
            ctx.manifolds_update()  # no step computed yet
        ctx.collide()
        ctx.manifolds_update()
        with pytest.raises(pk.PkError):
            ctx.manifolds_update()  # once per step
    finally:
        ctx.close()


def _feed(points):
    """One static pair (identity poses, so local = world), one new contact per step: (x, y, depth) on the plane z = 0."""
    M = oracle.Manifolds()
    key = np.array([(1 << 32) | 2], dtype=np.uint64)
    pos = np.zeros((3, 3))
    quat = np.tile([0.0, 0.0, 0.0, 1.0], (3, 1))
    for x, y, d in points:
        wa = np.array([x, y, 0.0])
        c = np.concatenate([[0.0, 0.0, 1.0], wa, wa + [0.0, 0.0, d], [d]])[None, :]
        M.step(key, np.array([1], np.uint8), c, pos, quat)
    mk, mc, mp = M.get()
    assert mc[0] == min(len(points), 4)
    return [(round(float(p[3]), 12), round(float(p[4]), 12), round(float(p[9]), 12)) for p in mp[0, : mc[0]]]


def test_add_reduce_known_answers_worked_out_by_hand():
    """manifold::add_reduce (collision_phases.h:139-198) on five points whose selection was worked out on paper from the
    reference's four rules — deepest; farthest from it; largest triangle with those two; farthest from the third —
    and its strict '>' comparisons (the first candidate wins a tie).  Independent of oracle/: the expected lists below
    were derived from the reference source, not produced by the restatement.

    Case A.  P0 (0, 0; 0.010)  P1 (0.1, 0.02; 0.030)  P2 (0.2, 0.08; 0.020)  P3 (0.3, 0.18; 0.015)  new P4 (0.4, 0.32; 0.040)
      deepest: P4.  |P−P4|²: P0 0.2624, P1 0.18, P2 0.0976, P3 0.0296 → P0.  edge0 = P0−P4 = (−0.4, −0.32); cross with
      P1−P4 = (−0.3, −0.30): 0.024, P2−P4 = (−0.2, −0.24): 0.032, P3−P4 = (−0.1, −0.14): 0.024 → P2.  Of P1, P3 the one
      farthest from P2: P1 0.0136, P3 0.02 → P3.  Kept, in this order: P4, P0, P2, P3.
    Case B (ties).  P0 (0, 0; 0.02)  P1 (0.1, 0; 0.02)  P2 (0, 0.3; 0.01)  P3 (0.05, 0.05; 0.01)  new P4 (0.3, 0; 0.02)
      deepest: P0 (P1, P4 are as deep, '>' keeps the first).  |P−P0|²: P1 0.01, P2 0.09, P3 0.005, P4 0.09 → P2 (P4 ties,
      comes later).  edge0 = P2−P0 = (0, 0.3); |cross|: P1 0.03, P3 0.015, P4 0.09 → P4.  Of P1, P3 the one farthest from
      P4: P1 0.04, P3 0.065 → P3.  Kept: P0, P2, P4, P3."""
    a = [(0.0, 0.0, 0.010), (0.1, 0.02, 0.030), (0.2, 0.08, 0.020), (0.3, 0.18, 0.015), (0.4, 0.32, 0.040)]
    assert _feed(a[:4]) == a[:4]  # below five points nothing is reduced or reordered (add_contact, :124-130)
    assert _feed(a) == [a[4], a[0], a[2], a[3]]
    b = [(0.0, 0.0, 0.02), (0.1, 0.0, 0.02), (0.0, 0.3, 0.01), (0.05, 0.05, 0.01), (0.3, 0.0, 0.02)]
    assert _feed(b) == [b[0], b[2], b[4], b[3]]


def test_merge_known_answers_worked_out_by_hand():
    """narrow_phase::calculate's merge (collision_phases.h:265-318) on cases whose outcome follows from the reference's
    constants by hand — distance2_eps = (5 mm)², contact_breaking_threshold = 5 cm, drift2_eps = 0.06 m² — for one pair
    (a = body 1, b = body 2, identity orientations, normal +z, world_b = world_a + depth·z as the reference's comment at
    :290 has it).  Expected values derived from the reference source, not from the restatement."""
    key = np.array([(1 << 32) | 2], dtype=np.uint64)
    quat = np.tile([0.0, 0.0, 0.0, 1.0], (3, 1))
    hit1, hit0 = np.array([1], np.uint8), np.array([0], np.uint8)

    def contact(x, y, d, shift_b=(0.0, 0.0, 0.0)):
        wa = np.array([x, y, 0.0])
        return np.concatenate([[0.0, 0.0, 1.0], wa, wa + [0.0, 0.0, d] + np.asarray(shift_b), [d]])[None, :]

    def start():
        M = oracle.Manifolds()
        began, ended = M.step(key, hit1, contact(0.0, 0.0, 0.01), np.zeros((3, 3)), quat)
        assert list(began) == [key[0]] and len(ended) == 0          # on_coll_beg (:309)
        M.set_impulses(np.array([[[1.5, 0.25, -0.75], [0, 0, 0], [0, 0, 0], [0, 0, 0]]]))
        return M

    # (1) a new contact 4 mm from the old one on body a (1.6e-5 < 2.5e-5 m²): the old point is dropped, the new one
    #     inherits its cached impulses (:272-281); result: one point, at the NEW position
    M = start()
    M.step(key, hit1, contact(0.004, 0.0, 0.012), np.zeros((3, 3)), quat)
    _, mc, mp = M.get()
    assert mc[0] == 1 and mp[0, 0, 3] == 0.004 and mp[0, 0, 9] == 0.012 and list(mp[0, 0, 10:13]) == [1.5, 0.25, -0.75]

    # (2) 6 mm away on both bodies (3.6e-5 > 2.5e-5): no match; the old point is re-projected (same poses: depth 0.01
    #     again) and kept with its impulses, the new one is appended with none (:283-303)
    M = start()
    M.step(key, hit1, contact(0.006, 0.0, 0.01), np.zeros((3, 3)), quat)
    _, mc, mp = M.get()
    assert mc[0] == 2 and mp[0, 0, 3] == 0.0 and mp[0, 1, 3] == 0.006
    assert list(mp[0, 0, 10:13]) == [1.5, 0.25, -0.75] and not mp[0, 1, 10:13].any()

    # (3) no new contact; b moved 5.5 cm away along the normal: depth = 0.01 − 0.055 = −0.045 > −0.05 → kept with that
    #     depth; 7 cm: −0.06 → dropped, the manifold empties and the collision ends (:300, :313-314)
    for dz, keep in ((0.055, True), (0.07, False)):
        M = start()
        pos = np.zeros((3, 3))
        pos[2, 2] = -dz
        began, ended = M.step(key, hit0, np.zeros((1, 10)), pos, quat)
        _, mc, mp = M.get()
        if keep:
            assert mc[0] == 1 and abs(mp[0, 0, 9] - (0.01 - dz)) < 1e-15 and len(ended) == 0
        else:
            assert M.count == 0 and list(ended) == [key[0]] and len(began) == 0

    # (4) no new contact; b slid 0.24 m along x: drift² = 0.0576 < 0.06 → kept, depth unchanged; 0.25 m: 0.0625 → dropped
    for dx, keep in ((0.24, True), (0.25, False)):
        M = start()
        pos = np.zeros((3, 3))
        pos[2, 0] = dx
        began, ended = M.step(key, hit0, np.zeros((1, 10)), pos, quat)
        _, mc, mp = M.get()
        if keep:
            assert mc[0] == 1 and abs(mp[0, 0, 9] - 0.01) < 1e-15 and len(ended) == 0
        else:
            assert M.count == 0 and list(ended) == [key[0]]
