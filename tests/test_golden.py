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
