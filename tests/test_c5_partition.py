"""BASELINE C5 (batched independent worlds) is split over ranks by worlds (reference: world_base::step per world,
core/world.h:234-242, src/world.cpp:20-56; no collective).  The split must be a partition, a world's content must not
depend on the number of ranks, and the ranks' results together must be the single-context result."""
import numpy as np
import pytest

import bench


@pytest.mark.parametrize("worlds,ranks", [(4096, 1), (4096, 2), (4096, 4), (4096, 8), (10, 4), (3, 8)])
def test_partition_covers_every_world_once(worlds, ranks):
    parts = bench.c5_partition(worlds, ranks)
    assert len(parts) == ranks
    seen = np.concatenate([np.arange(f, f + c) for f, c in parts]) if worlds else np.zeros(0)
    assert np.array_equal(seen, np.arange(worlds))
    counts = [c for _, c in parts]
    assert max(counts) - min(counts) <= 1


def test_world_content_independent_of_rank_count():
    base, pos, quat, sid, flags, wid = bench.c5_worlds(0, 6)
    per = base.n
    for first, cnt in bench.c5_partition(6, 4):
        if cnt == 0:
            continue
        b2, p2, q2, s2, f2, w2 = bench.c5_worlds(first, cnt)
        assert np.array_equal(p2, pos[first * per:(first + cnt) * per])
        assert np.array_equal(q2, quat[first * per:(first + cnt) * per])
        assert np.array_equal(w2, np.repeat(np.arange(cnt, dtype=np.uint32), per))


@pytest.mark.gpu
def test_ranks_together_give_the_single_context_result():
    """8 worlds in one context vs the same worlds as two 'ranks' of 5 + 3: pair keys and contact records of world k are
    the same whichever context it ran in (body ids shifted by the rank's first world)."""
    import physkit_b200 as pk

    def run(first, cnt):
        base, pos, quat, sid, flags, wid = bench.c5_worlds(first, cnt)
        n = len(pos)
        ctx = pk.Context(n, 14 * n, mode=pk.MODE_WORLD, max_shapes=8, max_contacts=7 * n, max_hull_vertices=64, num_worlds=cnt)
        ctx.add_shapes(base.shapes)
        ctx.resize(n)
        ctx.upload(pos, quat, None, sid, flags, wid)
        ctx.collide()
        p1 = pos.copy()
        p1[:, 1] -= 0.04
        ctx.update_pose(p1)
        ctx.collide()
        keys, con = ctx.pairs().copy(), ctx.contacts().copy()
        ctx.close()
        return base.n, keys, con

    per, keys_all, con_all = run(0, 8)
    assert len(keys_all) > 8 * 1000
    got_keys, got_con = [], []
    for first, cnt in ((0, 5), (5, 3)):
        _, k, c = run(first, cnt)
        off = np.uint64(first * per)
        shift = (off << np.uint64(32)) | off
        got_keys.append(k + shift)
        c = c.copy()
        c["key"] = c["key"] + shift
        got_con.append(c)
    assert np.array_equal(np.concatenate(got_keys), keys_all)
    assert np.array_equal(np.concatenate(got_con).view(np.uint8), con_all.view(np.uint8))
