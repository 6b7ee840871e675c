"""Golden vectors (tests/golden/, produced by tests/golden/make_golden.py from the KAT-pinned oracle):
the oracle must still reproduce them bit for bit (CPU), and so must the CUDA library through the C ABI (GPU)."""
import os

import numpy as np
import pytest

import oracle
from scenes import random_pairs_scene, scene_c3

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def _world_replay(step_fn, sc, steps):
    pos = sc.pos.copy()
    for step in range(steps):
        disp = np.full_like(pos, 0.01 * step)
        yield step, step_fn(pos, disp), pos
        pos = pos + 0.03


def test_oracle_reproduces_golden_world():
    g = np.load(os.path.join(GOLD, "world_c3_side6.npz"))
    sc = scene_c3(side=int(g["side"]))
    w = oracle.World(sc.shapes)

    def step(pos, disp):
        w.step(pos, sc.quat, disp, sc.shape_id, sc.flags)
        return w.pairs()

    for k, keys, used in _world_replay(step, sc, int(g["steps"])):
        assert np.array_equal(keys, g[f"keys{k}"]), f"pair set of step {k} moved"
    pa = (keys >> np.uint64(32)).astype(np.uint32)
    pb = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hit, out, _ = oracle.gjk_epa_pairs(sc.shapes, used, sc.quat, sc.shape_id, pa, pb)
    assert np.array_equal(hit, g["hit"])
    assert np.array_equal(_bits(out), _bits(g["contacts"]))


def test_oracle_reproduces_golden_pairs():
    g = np.load(os.path.join(GOLD, "pairs_mixed.npz"))
    sc, pa, pb = random_pairs_scene(int(g["n"]), int(g["seed"]))
    hit, out, _ = oracle.gjk_epa_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb)
    assert np.array_equal(hit, g["hit"])
    assert np.array_equal(_bits(out), _bits(g["contacts"]))
    assert 0.2 < hit.mean() < 0.8  # the fixture exercises both outcomes


@pytest.mark.gpu
def test_gpu_reproduces_golden_world():
    import physkit_b200 as pk
    from gpu_util import make_context

    g = np.load(os.path.join(GOLD, "world_c3_side6.npz"))
    sc = scene_c3(side=int(g["side"]))
    ctx = make_context(sc, max_pairs=50_000, mode=pk.MODE_WORLD)
    try:
        def step(pos, disp):
            ctx.upload(pos, sc.quat, disp, sc.shape_id, sc.flags)
            ctx.collide()
            return ctx.pairs()

        for k, keys, _ in _world_replay(step, sc, int(g["steps"])):
            assert np.array_equal(keys, g[f"keys{k}"]), f"pair set of step {k} differs from the golden one"
        con = ctx.contacts()
        m = g["hit"].astype(bool)
        assert np.array_equal(con["key"], keys[m])
        got = np.concatenate([con["normal"], con["world_a"], con["world_b"], con["depth"][:, None]], axis=1)
        assert np.array_equal(_bits(got), _bits(g["contacts"][m]))
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_reproduces_golden_pairs():
    import physkit_b200 as pk
    from gpu_util import contacts_equal_bitwise, make_context

    g = np.load(os.path.join(GOLD, "pairs_mixed.npz"))
    sc, pa, pb = random_pairs_scene(int(g["n"]), int(g["seed"]))
    ctx = make_context(sc, max_pairs=max(len(pa), 16), mode=pk.MODE_QUERY)
    try:
        hit, rec = ctx.gjk_epa_batch(pa, pb)
        contacts_equal_bitwise(rec, hit, g["hit"], g["contacts"])
    finally:
        ctx.close()


# ------------------------------------------------------------------ rows around the stage
def _rows_replay(g):
    import sys

    sys.path.insert(0, GOLD)
    from make_golden import downstream_state

    sc = scene_c3(side=int(g["side"]))
    return sc, downstream_state(sc.n)


def test_oracle_reproduces_golden_rows():
    g = np.load(os.path.join(GOLD, "rows_c3_side8.npz"))
    import sys

    sys.path.insert(0, GOLD)
    from make_golden import rows_c3

    got = rows_c3(int(g["side"]), int(g["steps"]))
    for k in ("man_keys", "man_counts", "row_keys", "row_points"):
        assert np.array_equal(got[k], g[k]), k
    assert np.array_equal(_bits(got["rows"]), _bits(g["rows"]))
    assert np.array_equal(got["rays"]["ray"], g["rays"]["ray"]) and np.array_equal(got["rays"]["body"], g["rays"]["body"])
    assert np.array_equal(_bits(got["rays"]["distance"]), _bits(g["rays"]["distance"]))
    assert np.array_equal(_bits(got["dyn_vel"]), _bits(g["dyn_vel"])) and np.array_equal(_bits(got["dyn_pos"]), _bits(g["dyn_pos"]))
    assert np.abs(got["dyn_quat"] - g["dyn_quat"]).max() < 1e-13  # sin / cos of the libm at hand
    assert len(g["rows"]) > 300 and len(g["rays"]) > 300


@pytest.mark.gpu
def test_gpu_reproduces_golden_rows():
    import physkit_b200 as pk
    from gpu_util import make_context

    g = np.load(os.path.join(GOLD, "rows_c3_side8.npz"))
    sc, (vel, w, mass, inertia, rest, fric, origins, dirs) = _rows_replay(g)
    ctx = make_context(sc, max_pairs=100_000, mode=pk.MODE_WORLD)
    ctx.manifolds_enable(10_000)
    ctx.dynamics_enable()
    ctx.dynamics_upload(vel, w, mass, inertia)
    ctx.material_upload(rest, fric)
    pos = sc.pos.copy()
    for step in range(int(g["steps"])):
        disp = np.full_like(pos, 0.01 * step)
        ctx.upload(pos, sc.quat, disp, sc.shape_id, sc.flags)
        ctx.collide()
        ctx.manifolds_update()
        pos = pos + 0.03
    man = ctx.manifolds()
    assert np.array_equal(man["key"], g["man_keys"]) and np.array_equal(man["count"], g["man_counts"])
    n = ctx.contact_rows_setup(1.0 / 60.0, 9.81)
    rows = ctx.contact_rows()
    assert n == len(g["rows"]) and np.array_equal(rows["key"], g["row_keys"]) and np.array_equal(rows["point"], g["row_points"])
    flat = np.concatenate([np.concatenate([rows[r]["J_v"], rows[r]["J_w_a"], rows[r]["J_w_b"], rows[r]["M_eff"][:, None], rows[r]["bias"][:, None]], axis=1)
                           for r in ("normal", "tangent1", "tangent2")] +
                          [rows["friction_coeff"][:, None], rows["inv_m_11"][:, None], rows["inv_m_12"][:, None], rows["inv_m_22"][:, None],
                           rows["accumulated"]], axis=1)
    assert np.array_equal(_bits(flat), _bits(g["rows"]))
    hits = ctx.raycast(origins, dirs, 6.0)
    assert np.array_equal(hits["ray"], g["rays"]["ray"]) and np.array_equal(hits["body"], g["rays"]["body"])
    assert np.array_equal(_bits(hits["distance"]), _bits(g["rays"]["distance"]))
    # integrator from the initial poses
    ctx.upload(sc.pos, sc.quat, np.zeros_like(sc.pos), sc.shape_id, sc.flags)
    ctx.dynamics_upload(vel, w, mass, inertia)
    for _ in range(3):
        ctx.integrate_velocities(1.0 / 60.0, (0.0, -9.81, 0.0))
        ctx.integrate_positions(1.0 / 60.0)
    p, q, v, om = ctx.dynamics_download(sc.n)
    assert np.array_equal(_bits(v), _bits(g["dyn_vel"])) and np.array_equal(_bits(p), _bits(g["dyn_pos"]))
    assert np.abs(q - g["dyn_quat"]).max() < 1e-13
    ctx.close()
