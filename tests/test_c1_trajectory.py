"""bench.py --workload c1: the prescribed drop of BASELINE C1 (ground + 10x10x10 unit boxes, spacing 1.2 m,
dt 1/60, 600 steps; the step order of src/world.cpp:22-55 hands vel*dt of the coming motion to update_node)."""
import numpy as np

import bench
import oracle


def test_c1_drop_is_a_fall_onto_a_settled_pile():
    sc, pos, disp = bench.c1_trajectory(603)
    assert sc.n == 1001 and pos.shape == (603, 1001, 3) and disp.shape == pos.shape
    assert np.array_equal(pos[0], sc.pos)                      # starts on the generator's lattice
    assert not disp[:, 0].any() and (pos[:, 0] == sc.pos[0]).all()  # the ground never moves
    assert (disp[:, :, [0, 2]] == 0).all() and (disp[:, :, 1] <= 0).all()  # straight down
    assert np.allclose(pos[1:], pos[:-1] + disp[:-1], rtol=0, atol=1e-12)  # disp = the motion of the coming step
    layer = np.round((sc.pos[1:, 1] - 1.0) / 1.2)
    assert np.allclose(pos[-1, 1:, 1], 0.5 + 0.98 * layer)     # settled: layers 0.98 m apart, 2 cm of resting overlap
    assert not disp[-1].any()
    # free fall until then: the first step moves every box by g·dt²
    assert np.allclose(disp[0, 1:, 1], -9.81 / 3600.0)


def test_c1_pairs_grow_from_the_first_step_quirk_to_resting_contacts():
    sc, pos, disp = bench.c1_trajectory(200)
    w = oracle.World(sc.shapes)
    counts = []
    for k in range(200):
        w.step(pos[k], sc.quat, disp[k], sc.shape_id, sc.flags)
        counts.append(len(w.pairs()))
    assert counts[0] == 0            # collision_phases.h:342-346: nothing has moved before the first step
    assert counts[1] > 1000 and counts[-1] >= counts[1]
    keys = w.pairs()
    hit, _, _ = oracle.gjk_epa_pairs(sc.shapes, pos[199], sc.quat, sc.shape_id, (keys >> np.uint64(32)).astype(np.uint32),
                                     (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32), nthreads=8)
    assert hit.sum() >= 900          # every box of the settled pile rests on the one below (or on the ground)
