"""The FP32 miss filter of the GJK stage (physkit_b200/csrc/pk_gjk_filter.cuh), run on the host through tests/emul.py,
against the oracle: a pair the filter drops must be a miss of the reference's gjk_collision — on random pairs of all
kinds, on C3 / C4 pairs and on pairs that graze each other at gaps from 1e-12 of their size up — and on the analytic
kinds (spheres, boxes) it must drop practically every miss."""
import numpy as np
import pytest

import emul
import oracle
from scenes import near_touching_scene, random_pairs_scene, scene_c3, scene_c4

pytestmark = pytest.mark.skipif(not emul.available(), reason="CUDA headers not installed")


def _check(sc, pa, pb, iters=2):
    hit, _, _ = oracle.gjk_epa_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb, stats=True, nthreads=8)
    drop = emul.filter_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb, iters)
    wrong = np.nonzero((drop == 1) & (hit == 1))[0]
    assert len(wrong) == 0, f"filter dropped hits: pairs {wrong[:10]}"
    miss = hit == 0
    return int(miss.sum()), int((miss & (drop == 0)).sum())


@pytest.mark.parametrize("seed,far", [(31, 1e2), (32, 1e4), (33, 1e6)])
def test_filter_never_drops_a_grazing_hit(seed, far):
    misses, kept = _check(*near_touching_scene(20_000, seed, far=far))
    assert kept < 0.5 * misses  # (grazing misses inside the margin are kept; the clear ones are dropped)


def test_filter_on_random_pairs_of_all_kinds():
    misses, kept = _check(*random_pairs_scene(20_000, 11))
    assert kept < 0.1 * misses


def test_filter_is_complete_on_c3_pairs():
    """Spheres and boxes: sphere–sphere, sphere–box distance and the 15 box–box axes decide every pair outside the margin."""
    sc = scene_c3(side=16)
    boxes = oracle.bounds(sc.shapes, sc.pos, sc.quat, sc.shape_id)
    boxes[:, :3] -= 0.1
    boxes[:, 3:] += 0.1
    keys = oracle.query_pairs(boxes)
    misses, kept = _check(sc, (keys >> np.uint64(32)).astype(np.uint32), (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32))
    assert misses > 40_000 and kept < 0.001 * misses


@pytest.mark.parametrize("iters", [0, 2, 8])
def test_filter_on_c4_hull_pairs(iters):
    misses, kept = _check(*scene_c4(n_pairs=6_000, n_hulls=64), iters=iters)
    assert kept < 0.08 * misses


def test_filter_leaves_non_unit_quaternions_to_the_exact_path():
    """A quaternion that is not unit makes the reference's box a parallelepiped (lin_alg.h:493-499 is then not a rotation);
    the box tests assume a frame, so such bodies are not filtered at all."""
    sc, pa, pb = random_pairs_scene(2_000, 5, kinds=("obb",), spread=6.0)
    sc.quat[0::2] *= 1.01
    drop = emul.filter_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb)
    assert drop.sum() == 0
    sc.quat[0::2] /= 1.01
    assert emul.filter_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb).sum() > 1_000


def _adversarial(seed, mutate):
    sc, pa, pb = random_pairs_scene(6_000, seed, kinds=("obb", "sphere", "hull", "aabb"))
    mutate(sc)
    hit, _, _ = oracle.gjk_epa_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb, stats=True, nthreads=8)
    drop = emul.filter_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb)
    wrong = np.nonzero((drop == 1) & (hit == 1))[0]
    assert len(wrong) == 0, f"filter dropped hits: pairs {wrong[:10]}"
    return hit, drop


def test_filter_with_non_finite_poses():
    """NaN / inf positions and quaternions: whatever the reference's iteration makes of them (it can end in "hit": every
    comparison against NaN is false, and an all-false tetrahedron test encloses the origin), the filter must not have
    an opinion — the margin is NaN, no gap exceeds it."""

    def mutate(sc):
        sc.pos[0::7, 0] = np.nan
        sc.pos[3::11, 2] = np.inf
        sc.quat[5::13, 1] = np.nan

    kinds = {}

    def mutate_and_note(sc):
        mutate(sc)
        kinds["k"] = [s[0] for s in sc.shapes]

    hit, drop = _adversarial(41, mutate_and_note)
    kind = np.array(kinds["k"])  # (shape i belongs to body i in this scene)
    n = len(kind)
    bad = np.zeros(n, bool)
    idx = np.arange(n)
    bad |= ((idx % 7 == 0) | (idx % 11 == 3)) & (kind != "aabb")  # a world box does not look at the body's position,
    bad |= (idx % 13 == 5) & ((kind == "obb") | (kind == "hull"))  # nor a sphere at its orientation
    bad_pair = bad[0::2] | bad[1::2]
    assert bad_pair.sum() > 1_000 and drop[bad_pair].sum() == 0


def test_filter_with_negative_and_zero_extents():
    """support() picks among the corners (±hx, ±hy, ±hz) / the points p ± r·d whatever the signs stored, and a shape of
    size zero is a point: the filter works with |h| and |r| and keeps its absolute floor of 2e-6."""

    def mutate(sc):
        for i, s in enumerate(sc.shapes):
            if s[0] == "obb" and i % 3 == 0:
                sc.shapes[i] = ("obb", -np.asarray(s[1]))
            elif s[0] == "sphere" and i % 3 == 0:
                sc.shapes[i] = ("sphere", -s[1])
            elif s[0] == "sphere" and i % 3 == 1:
                sc.shapes[i] = ("sphere", 0.0)
            elif s[0] == "obb" and i % 3 == 1:
                sc.shapes[i] = ("obb", np.zeros(3))

    hit, drop = _adversarial(42, mutate)
    assert drop.sum() > 1_000  # still filtering


def test_filter_far_from_the_origin_and_at_extreme_sizes():
    """Coordinates of 1e9 (FP32 could not tell the bodies of a pair apart: the centres are subtracted in FP64 first) and
    scenes scaled by 1e-6 and 1e6."""
    for k, scale in enumerate((1e-6, 1.0, 1e6)):

        def mutate(sc, scale=scale):
            sc.pos *= scale
            sc.pos += 1e9 * scale
            for i, s in enumerate(sc.shapes):
                if s[0] == "sphere":
                    sc.shapes[i] = ("sphere", s[1] * scale)
                elif s[0] == "aabb":
                    sc.shapes[i] = ("aabb", sc.pos[i] - (np.asarray(s[2]) - np.asarray(s[1])) * 0.5 * scale,
                                    sc.pos[i] + (np.asarray(s[2]) - np.asarray(s[1])) * 0.5 * scale)
                else:
                    sc.shapes[i] = (s[0], np.asarray(s[1]) * scale)

        hit, drop = _adversarial(43 + k, mutate)
        if scale >= 1.0:
            assert drop.sum() > 1_000
