"""pk_gjk_distance_batch's per-pair routine (gjk_distance_pair, physkit_b200/csrc/pk_distance.cuh), run on the host through
tests/emul.py, checked three independent ways:

* against the oracle's brute-force distance over all vertex / edge / triangle combinations (oracle.distance_brute_pairs);
* by certificate, in numpy: the witness points lie in their bodies, are `distance` apart, and the slab between the
  supporting planes perpendicular to their difference is `distance` wide (no two points of the bodies can be closer);
* against the reference's boolean query (oracle.gjk_epa_pairs = gjk_collision, src/collision.cpp:165-189): separated
  pairs are misses, the others hits, outside its 1e-6 m margin.

The reference has no distance query, so there are no vectors of its own to pin this on."""
import numpy as np
import pytest
from scipy.spatial import ConvexHull

import emul
import oracle
from scenes import _quat_matrix, near_touching_scene, random_pairs_scene

pytestmark = pytest.mark.skipif(not emul.available(), reason="CUDA headers not installed")


def world_core(sc, i):
    """→ (vertices[n,3] of body i's core in the world frame, margin)."""
    spec = sc.shapes[sc.shape_id[i]]
    R = _quat_matrix(sc.quat[i])
    if spec[0] == "sphere":
        return sc.pos[i][None, :].copy(), float(spec[1])
    if spec[0] == "aabb":
        lo, hi = np.asarray(spec[1], float), np.asarray(spec[2], float)
        return np.array([[hi[0] if j & 1 else lo[0], hi[1] if j & 2 else lo[1], hi[2] if j & 4 else lo[2]] for j in range(8)]), 0.0
    if spec[0] == "obb":
        h = np.asarray(spec[1], float)
        loc = np.array([[h[0] if j & 1 else -h[0], h[1] if j & 2 else -h[1], h[2] if j & 4 else -h[2]] for j in range(8)])
    else:
        loc = np.asarray(spec[1], float)
    return loc @ R.T + sc.pos[i], 0.0


def in_body(verts, margin, p, tol):
    if len(verts) == 1:
        return np.linalg.norm(p - verts[0]) <= margin + tol
    try:
        eq = ConvexHull(verts).equations
    except Exception:  # flat hull: QHull refuses; covered by the brute-force comparison
        return True
    return bool((eq[:, :3] @ p + eq[:, 3] <= tol).all())


def check_certificates(sc, pa, pb, sep, rec, sample):
    for k in sample:
        if not sep[k]:
            assert rec["distance"][k] == 0.0 and not rec["point_a"][k].any() and not rec["point_b"][k].any()
            continue
        VA, ra = world_core(sc, pa[k])
        VB, rb = world_core(sc, pb[k])
        a, b, d = rec["point_a"][k], rec["point_b"][k], rec["distance"][k]
        scale = max(np.abs(VA).max(), np.abs(VB).max(), ra, rb, 1e-300)
        tol = 1e-10 * scale
        assert d > 0.0
        assert abs(np.linalg.norm(a - b) - d) <= tol, (k, np.linalg.norm(a - b), d)
        assert in_body(VA, ra, a, 10 * tol) and in_body(VB, rb, b, 10 * tol), k
        if d < 1e-5 * scale:
            continue  # the direction of a difference of 1e-5 of the coordinates is not known to 1e-10 (brute force covers these)
        n = (b - a) / np.linalg.norm(b - a)
        slab = ((VB @ n).min() - rb) - ((VA @ n).max() + ra)  # no point of B is closer to A than this along n
        assert slab >= d - tol, (k, slab, d)


def check_against_brute(sc, pa, pb, sep, rec, hit, max_verts=32):
    brute = oracle.distance_brute_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb, max_verts=max_verts)
    scale = np.maximum(np.abs(sc.pos[pa]).max(axis=1), np.abs(sc.pos[pb]).max(axis=1)) + 1.0
    have = ~np.isnan(brute)
    s = have & (sep == 1)
    assert s.sum() > 0.2 * len(pa)
    err = np.abs(rec["distance"][s] - brute[s])
    assert (err <= 1e-10 * scale[s]).all(), f"worst {err.max():.3e} at {np.nonzero(s)[0][err.argmax()]}"
    # not separated ⇒ the bodies do touch or overlap.  The reference's boolean query cannot be the judge at every scale
    # (its degeneracy thresholds are absolute, src/collision.cpp:25,57,69: below sizes of 1e-2 m it misses overlapping
    # pairs), so: the origin must lie in the hull of the vertex differences, or within the margins of it.
    ns = np.nonzero(have & (sep == 0))[0]
    for k in ns[:: max(1, len(ns) // 150)]:
        VA, ra = world_core(sc, pa[k])
        VB, rb = world_core(sc, pb[k])
        M = (VA[:, None, :] - VB[None, :, :]).reshape(-1, 3)
        size = np.abs(M).max()
        if len(M) < 4:
            assert np.linalg.norm(M[0]) <= ra + rb + 1e-9 * size
            continue
        try:
            off = ConvexHull(M).equations[:, 3].max()  # > 0: the origin is outside, at least this far
        except Exception:
            continue
        assert off <= ra + rb + 1e-7 * size, (k, off, ra + rb)
    return brute


def run(sc, pa, pb):
    sep, rec = emul.distance_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb)
    hit, _, _ = oracle.gjk_epa_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb, nthreads=8)
    assert (rec["key"] == (pa.astype(np.uint64) << np.uint64(32) | pb.astype(np.uint64))).all()
    # agreement with gjk_collision outside its margin: a separated pair with more than 1e-6 of clearance is a miss,
    # and a pair reported as touching / overlapping is a hit unless it grazes
    clear = (sep == 1) & (rec["distance"] > 1e-6)
    assert (hit[clear] == 0).all(), np.nonzero(clear & (hit == 1))[0][:10]
    return sep, rec, hit


@pytest.mark.parametrize("seed,spread", [(3, 1.2), (4, 2.5), (5, 6.0)])
def test_distance_on_random_pairs_of_all_kinds(seed, spread):
    sc, pa, pb = random_pairs_scene(3_000, seed, spread=spread)
    sep, rec, hit = run(sc, pa, pb)
    assert 0.1 < sep.mean() < 0.999 or spread > 5
    check_against_brute(sc, pa, pb, sep, rec, hit)
    check_certificates(sc, pa, pb, sep, rec, range(0, len(pa), 7))
    touching = sep == 0
    assert (hit[touching] == 1).mean() > 0.99  # (the rest graze within 1e-6: checked against brute force above)


@pytest.mark.parametrize("kinds", [("obb",), ("sphere",), ("sphere", "obb"), ("hull",), ("aabb", "obb")])
def test_distance_per_kind(kinds):
    sc, pa, pb = random_pairs_scene(1_500, 17, kinds=kinds, spread=2.0)
    sep, rec, hit = run(sc, pa, pb)
    check_against_brute(sc, pa, pb, sep, rec, hit)
    check_certificates(sc, pa, pb, sep, rec, range(0, len(pa), 5))


def test_sphere_sphere_is_exact():
    sc, pa, pb = random_pairs_scene(2_000, 23, kinds=("sphere",), spread=3.0)
    sep, rec, _ = run(sc, pa, pb)
    r = np.array([sc.shapes[i][1] for i in range(sc.n)])
    want = np.linalg.norm(sc.pos[pa] - sc.pos[pb], axis=1) - r[pa] - r[pb]
    assert ((want > 1e-12) == (sep == 1)).all()
    s = sep == 1
    assert np.abs(rec["distance"][s] - want[s]).max() <= 1e-14 * 10


@pytest.mark.parametrize("seed,far", [(31, 1e2), (32, 1e4)])
def test_distance_of_grazing_pairs(seed, far):
    """Gaps of 0, ±1e-12 … ±1e-1 of the pair's size along a known axis, three size scales, far from the origin: the
    distance is at least the slab the generator left (exactly it when the axis is a face normal of a box)."""
    sc, pa, pb = near_touching_scene(4_000, seed, far=far)
    sep, rec, hit = run(sc, pa, pb)
    check_certificates(sc, pa, pb, sep, rec, range(0, len(pa), 9))
    small = np.array([sc.shapes[i][0] != "hull" or len(sc.shapes[i][1]) <= 32 for i in range(sc.n)])
    m = small[pa] & small[pb]
    check_against_brute(sc, pa[m], pb[m], sep[m], rec[m], hit[m])
    assert (sep == 1).sum() > 1_000


def test_big_hulls_and_scales():
    sc, pa, pb = near_touching_scene(1_500, 41, kinds=("bighull", "obb", "sphere"), scales=(1e-3, 1.0, 1e3), far=10.0)
    sep, rec, _ = run(sc, pa, pb)
    assert (sep == 1).sum() > 400
    check_certificates(sc, pa, pb, sep, rec, range(0, len(pa), 3))


def test_degenerate_inputs():
    """Coincident bodies, a body against itself, flat and needle boxes, non-finite poses: never a crash, never a distance
    for a pair that is not separated."""
    shapes = [("obb", (0.5, 0.5, 0.5)), ("obb", (1.0, 1e-9, 1.0)), ("obb", (1e-9, 1e-9, 2.0)), ("sphere", 0.0), ("sphere", 1.0)]
    pos = np.array([[0, 0, 0], [0, 0, 0], [0, 3.0, 0], [5.0, 0, 0], [0, 0, 4.0], [np.nan, 0, 0], [0, np.inf, 0]], dtype=float)
    quat = np.tile([0, 0, 0, 1.0], (len(pos), 1))
    sid = np.array([0, 0, 1, 2, 3, 4, 4], dtype=np.uint32)
    sc_pa = np.array([0, 0, 0, 0, 0, 2, 3, 0, 0, 4], dtype=np.uint32)
    sc_pb = np.array([0, 1, 2, 3, 4, 3, 4, 5, 6, 4], dtype=np.uint32)
    sep, rec = emul.distance_pairs(shapes, pos, quat, sid, sc_pa, sc_pb)
    assert list(sep) == [0, 0, 1, 1, 1, 1, 1, 0, 0, 0]
    assert rec["distance"][2] == pytest.approx(3.0 - 0.5 - 1e-9, abs=1e-12)      # cube – flat plate
    assert rec["distance"][3] == pytest.approx(5.0 - 0.5 - 1e-9, abs=1e-12)      # cube – needle
    assert rec["distance"][4] == pytest.approx(4.0 - 0.5, abs=1e-12)             # cube – point (sphere of radius 0)
    assert rec["distance"][5] == pytest.approx(np.hypot(4.0 - 1e-9, 3.0 - 2e-9), rel=1e-9)  # plate – needle
    assert (rec["distance"][sep == 0] == 0.0).all() and np.isfinite(rec["distance"]).all()


def test_distance_is_symmetric_and_translation_invariant():
    """d(a, b) = d(b, a) with the witness points swapped, and moving both bodies together changes nothing beyond the
    rounding of the moved coordinates (the iteration runs in a frame attached to the first body)."""
    sc, pa, pb = random_pairs_scene(2_000, 77, spread=2.5)
    sep, rec = emul.distance_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb)
    sep2, rec2 = emul.distance_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pb, pa)
    assert np.array_equal(sep, sep2)
    s = sep == 1
    assert np.abs(rec["distance"][s] - rec2["distance"][s]).max() <= 1e-12
    # (the closest points need not be unique — parallel faces, edges —, so only each answer's own pair is compared)
    assert np.abs(np.linalg.norm(rec2["point_a"][s] - rec2["point_b"][s], axis=1) - rec["distance"][s]).max() <= 1e-12
    shift = np.array([1234.5, -987.25, 4321.125])
    moved = [("aabb", np.asarray(spec[1]) + shift, np.asarray(spec[2]) + shift) if spec[0] == "aabb" else spec for spec in sc.shapes]
    sep3, rec3 = emul.distance_pairs(moved, sc.pos + shift, sc.quat, sc.shape_id, pa, pb)
    assert np.array_equal(sep, sep3)
    assert np.abs(rec["distance"][s] - rec3["distance"][s]).max() <= 1e-11
    assert np.abs(np.linalg.norm(rec3["point_a"][s] - rec3["point_b"][s], axis=1) - rec["distance"][s]).max() <= 1e-11
