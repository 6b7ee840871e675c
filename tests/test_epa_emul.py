"""The EPA kernels' control flow, checked in the container without a GPU.

tests/emul.py compiles the kernels' own source (pk_epa_coop.cuh and what it builds on) with g++ and runs one OS thread
per CUDA thread (tests/cpp/simt_host.h).  What this pins down before any GPU time is spent: the split of an iteration
into the per-pair part and the per-edge part dealt out over the warp, the restart of tied pairs in HEAP mode, the hand-back to epa_kernel, the
parked results / refill protocol, the pinned-buffer copy — all bit for bit against the oracle.  A kernel whose lanes
disagree about a warp-wide vote dead-locks here (pytest time-out) instead of on the GPU box.  The same scenes run through
the real kernels in tests/test_gpu_narrowphase.py."""
import numpy as np
import pytest

import emul
import oracle
from scenes import IDENT, Scene, SplitMix64, random_pairs_scene, scene_c3, scene_c4, box_vertices, sphere_vertices

pytestmark = [pytest.mark.skipif(not emul.available(), reason="CUDA headers not installed"), pytest.mark.timeout(600)]


def _check(sc, pa, pb, mirror=False, capacity=None, arrival=2):
    hit, out, stats = emul.gjk_epa_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb, capacity=capacity, mirror=mirror, arrival=arrival)
    hit_ref, out_ref, _ = oracle.gjk_epa_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb, nthreads=8)
    assert np.array_equal(hit, hit_ref), f"hit flags differ at {np.nonzero(hit != hit_ref)[0][:10]}"
    m = hit_ref.astype(bool)
    got = np.concatenate([out["normal"], out["world_a"], out["world_b"], out["depth"][:, None]], axis=1)
    a = got[m].view(np.uint64)
    b = np.ascontiguousarray(out_ref[m]).view(np.uint64)
    bad = np.nonzero((a != b).any(axis=1))[0]
    assert len(bad) == 0, f"{len(bad)} of {m.sum()} contacts differ; first pair {np.nonzero(m)[0][bad[0]]}: {got[m][bad[0]]} vs {out_ref[m][bad[0]]}"
    return hit, stats


def _lattice_pairs(sc, reach):
    """candidate pairs of a small lattice scene: all (i<j) closer than `reach`"""
    d = sc.pos[:, None, :] - sc.pos[None, :, :]
    close = (np.abs(d) < reach).all(axis=2)
    i, j = np.nonzero(np.triu(close, 1))
    return i.astype(np.uint32), j.astype(np.uint32)


@pytest.mark.parametrize("mirror", [False, True])
def test_c3_mix_spheres_boxes(mirror):
    """BASELINE C3's mix: sphere–sphere, sphere–box (SCAN instance) and box–box (HEAP instance) pairs."""
    sc = scene_c3(side=9)
    pa, pb = _lattice_pairs(sc, 1.0)
    hit, stats = _check(sc, pa, pb, mirror=mirror)
    assert hit.sum() > 300 and stats["class0"] > 50 and stats["class2"] > 10 and stats["class1"] > 300
    assert stats["valid"] + stats["dropped"] <= stats["gjk_hits"]


def test_all_shape_kinds():
    sc, pa, pb = random_pairs_scene(1500, 21)
    hit, stats = _check(sc, pa, pb)
    assert 0.2 < hit.mean() < 0.9


def test_hull_pairs_c4():
    sc, pa, pb = scene_c4(n_pairs=400, n_hulls=16)
    hit, _ = _check(sc, pa, pb)
    assert hit.sum() > 100


def test_mesh_boxes_and_mesh_spheres_tie_heavy():
    """8-vertex mesh boxes stacked face to face (coplanar faces: distance ties are the rule) and 482-vertex mesh spheres."""
    rng = SplitMix64(77)
    n = 120
    shapes = [("hull", box_vertices((0.5, 0.5, 0.5))), ("hull", sphere_vertices(0.5)), ("obb", (0.5, 0.5, 0.5)), ("sphere", 0.5)]
    pos = np.zeros((2 * n, 3))
    pos[0::2] = rng.uniform(-3, 3, n, 3)
    off = rng.uniform(-0.9, 0.9, n, 3)
    off[: n // 2] = np.round(off[: n // 2] * 2) / 2  # axis-aligned offsets by multiples of 0.5: exact ties
    pos[1::2] = pos[0::2] + off
    quat = np.tile(np.array([0.0, 0.0, 0.0, 1.0]), (2 * n, 1))
    quat[n:] = rng.quats(n)
    sid = rng.randint(2 * n, 4).astype(np.uint32)
    sc = Scene(shapes, pos, quat, sid)
    pa = np.arange(0, 2 * n, 2, dtype=np.uint32)
    pb = pa + 1
    _check(sc, pa, pb)


def test_more_hits_than_contact_records():
    """GJK hits beyond the contact capacity are dropped and counted, nothing is written past the arrays
    (out_slot was unchecked).  The records that fit are still the oracle's."""
    sc = scene_c3(side=6)
    pa, pb = _lattice_pairs(sc, 1.0)
    hit_ref, out_ref, _ = oracle.gjk_epa_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb)
    cap = int(hit_ref.sum()) // 2
    hit, out, stats = emul.gjk_epa_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb, capacity=cap)
    assert stats["gjk_hits"] == hit_ref.sum() and stats["dropped"] > 0
    rank = np.cumsum(hit_ref) - hit_ref
    assert not hit[(rank >= cap) & (hit_ref == 1)].any()
    assert hit.sum() > 0 and stats["valid"] == hit.sum()
    got = np.concatenate([out["normal"], out["world_a"], out["world_b"], out["depth"][:, None]], axis=1)
    m = hit.astype(bool)
    assert np.array_equal(got[m].view(np.uint64), np.ascontiguousarray(out_ref[m]).view(np.uint64))


def test_reference_kat_cases_in_one_batch():
    """All GJK / EPA cases of the reference's tests (tests/kat_cases.py) as one batch: same-box and containment cases
    reach pad_simplex, i.e. the second hand-back (HEAP → epa_kernel)."""
    from kat_cases import EPA_CASES, GJK_CASES

    shapes, pos, quat = [], [], []
    for c in list(GJK_CASES) + list(EPA_CASES):
        for spec, p, q in (c[1], c[2]):
            shapes.append(spec)
            pos.append(p)
            quat.append(q)
    # the first support point is the origin (collision.cpp:174): GJK returns a one-point simplex, EPA pads it
    for r, gap in ((1.0, 2.0), (0.5, 1.0), (0.25, 0.5)):
        shapes += [("sphere", r), ("sphere", r)]
        pos += [(0.0, 0.0, 0.0), (gap, 0.0, 0.0)]
        quat += [IDENT, IDENT]
    n = len(shapes) // 2
    sc = Scene(shapes, np.array(pos), np.array(quat), np.arange(2 * n))
    pa = np.arange(0, 2 * n, 2, dtype=np.uint32)
    hit, stats = _check(sc, pa, pa + 1)
    assert stats["handed_to_epa_kernel"] > 0


def test_sphere_pairs_with_exact_ties_restart_in_heap_mode():
    """Axis-aligned equal spheres: mirror-symmetric polytopes whose equidistant faces only the heap's history can order.
    Most pairs are started again in HEAP mode by their lane; two blocks: the second finds the list drained."""
    rng = SplitMix64(5)
    n = 600
    shapes = [("sphere", 0.5), ("sphere", 0.3)]
    pos = np.zeros((2 * n, 3))
    pos[0::2] = rng.uniform(-3, 3, n, 3)
    pos[1::2] = pos[0::2] + rng.uniform(-0.4, 0.4, n, 3)
    quat = np.tile([0, 0, 0, 1.0], (2 * n, 1))
    sid = rng.randint(2 * n, 2).astype(np.uint32)
    sc = Scene(shapes, pos, quat, sid)
    pa = np.arange(0, 2 * n, 2, dtype=np.uint32)
    hit, out, stats = emul.gjk_epa_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pa + 1, nblocks=2)
    assert stats["restarted_in_heap_mode"] > 100 and stats["valid"] == hit.sum()
    _check(sc, pa, pa + 1)


# ---- the whole stage as run_narrowphase launches it: gjk_filter_kernel / gjk_prefilter_kernel → gjk_kernel → flag scan →
# ---- epa_order_kernel → epa_init_kernel → epa_coop_kernel → epa_kernel (arrival=3: no host-side stand-in for the front)
def test_whole_stage_kernels_c3_mix():
    """Spheres and boxes: misses settled by the FP32 filter kernel, survivors listed per shape-kind class, hits appended
    in the order the threads happen to run, contact slots from the scan, EPA order from epa_order_kernel."""
    sc = scene_c3(side=9)
    pa, pb = _lattice_pairs(sc, 1.0)
    hit, stats = _check(sc, pa, pb, arrival=3)
    assert hit.sum() > 300 and stats["class0"] > 50 and stats["class2"] > 10 and stats["class1"] > 300
    assert stats["valid"] == hit.sum()


def test_whole_stage_kernels_all_shape_kinds_and_hull_contexts():
    """All kinds (the hull context takes gjk_prefilter_kernel<true, true> and carries two support points to gjk_kernel)."""
    sc, pa, pb = random_pairs_scene(1500, 21)
    hit, _ = _check(sc, pa, pb, arrival=3)
    assert 0.2 < hit.mean() < 0.9
    sc, pa, pb = scene_c4(n_pairs=400, n_hulls=16)
    hit, _ = _check(sc, pa, pb, arrival=3)
    assert hit.sum() > 100


def test_whole_stage_kernels_more_hits_than_contact_records():
    """GJK hits beyond the contact capacity are counted, not written (the contact slot was unchecked)."""
    sc = scene_c3(side=7)
    pa, pb = _lattice_pairs(sc, 1.0)
    full, _, _ = emul.gjk_epa_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb, arrival=3)
    cap = int(full.sum()) // 2
    hit, out, stats = emul.gjk_epa_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb, capacity=cap, arrival=3)
    assert stats["gjk_hits"] == full.sum() and stats["valid"] <= cap and hit.sum() == stats["valid"]
    # the records that were written are the oracle's
    _, out_ref, _ = oracle.gjk_epa_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb, nthreads=8)
    m = hit.astype(bool)
    got = np.concatenate([out["normal"], out["world_a"], out["world_b"], out["depth"][:, None]], axis=1)
    assert np.array_equal(got[m].view(np.uint64), np.ascontiguousarray(out_ref[m]).view(np.uint64))
