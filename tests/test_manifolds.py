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
