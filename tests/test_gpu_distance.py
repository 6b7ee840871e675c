"""GPU: pk_gjk_distance_batch (gjk_distance_kernel, physkit_b200/csrc/pk_distance.cuh) through the C ABI.

The reference has no distance query, so the checkers are the oracle's brute-force distance over all
feature pairs and separation certificates (tests/test_gjk_distance_host.py), plus — bit for bit — the host run of the
kernel's own per-pair source, and the reference's boolean query for the hit / miss verdict outside its 1e-6 m margin."""
import numpy as np
import pytest

import emul
import oracle
from gpu_util import make_context
from scenes import near_touching_scene, random_pairs_scene, scene_c3, scene_c4
from test_gjk_distance_host import check_against_brute, check_certificates

pytestmark = pytest.mark.gpu


def _gpu(sc, pa, pb):
    ctx = make_context(sc, max(len(pa), 16))
    try:
        return ctx.gjk_distance_batch(pa, pb)
    finally:
        ctx.close()


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def _same_as_host(sc, pa, pb, sep, rec):
    if not emul.available():
        return
    hsep, hrec = emul.distance_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb)
    assert np.array_equal(sep, hsep)
    for f in ("key", "distance", "point_a", "point_b"):
        assert np.array_equal(_bits(rec[f]), _bits(hrec[f])), f


@pytest.mark.parametrize("seed,spread", [(3, 1.2), (5, 6.0)])
def test_distance_random_pairs_all_kinds(seed, spread):
    sc, pa, pb = random_pairs_scene(3_000, seed, spread=spread)
    sep, rec = _gpu(sc, pa, pb)
    hit, _, _ = oracle.gjk_epa_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb, nthreads=8)
    assert (hit[(sep == 1) & (rec["distance"] > 1e-6)] == 0).all()
    assert (hit[sep == 0] == 1).mean() > 0.99
    check_against_brute(sc, pa, pb, sep, rec, hit)
    check_certificates(sc, pa, pb, sep, rec, range(0, len(pa), 7))
    _same_as_host(sc, pa, pb, sep, rec)


def test_distance_grazing_pairs_with_big_hulls():
    """The instance for contexts with many-vertex hulls (support<true>: float-prefiltered scan, cooperative exact pass)
    returns the same vertices as the plain scan the host run uses: bit-identical records."""
    sc, pa, pb = near_touching_scene(4_000, 31, far=1e2)
    sep, rec = _gpu(sc, pa, pb)
    assert (sep == 1).sum() > 1_000
    check_certificates(sc, pa, pb, sep, rec, range(0, len(pa), 9))
    _same_as_host(sc, pa, pb, sep, rec)


def test_distance_c4_hull_pairs():
    sc, pa, pb = scene_c4(n_pairs=20_000, n_hulls=64)
    sep, rec = _gpu(sc, pa, pb)
    hit, _, _ = oracle.gjk_epa_pairs(sc.shapes, sc.pos, sc.quat, sc.shape_id, pa, pb, nthreads=8)
    assert (hit[(sep == 1) & (rec["distance"] > 1e-6)] == 0).all()
    assert (hit[sep == 0] == 1).mean() > 0.99
    check_certificates(sc, pa, pb, sep, rec, range(0, len(pa), 97))
    _same_as_host(sc, pa[:4_000], pb[:4_000], sep[:4_000], rec[:4_000])


def test_distance_leaves_the_step_alone_and_runs_resident():
    """The query uses no buffer of the step: pairs and contacts of the last pk_collide are still there afterwards; the
    device-pointer form gives the same records and a device time."""
    sc = scene_c3(side=24)
    ctx = make_context(sc, 1 << 20, mode=__import__("physkit_b200").MODE_WORLD)
    try:
        ctx.collide_resident()  # (the reference's first step reports no pair: broad_phase's first-step quirk)
        ctx.update_pose(sc.pos + 0.05, None, None)
        ctx.collide_resident()
        ctx.fetch()
        keys, contacts = ctx.pairs(), ctx.contacts()
        assert len(contacts) > 1_000
        pa = (keys >> np.uint64(32)).astype(np.uint32)
        pb = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        sep, rec = ctx.gjk_distance_batch(pa, pb)
        touching = np.isin(keys, contacts["key"])
        # every contact of the step is a pair without a distance (or one that grazes inside gjk_collision's margin)
        assert (rec["distance"][touching] < 1e-6).all() and (sep[touching] == 0).mean() > 0.999
        assert (sep == 1).sum() > 0.5 * len(keys)
        ctx.fetch()  # the device-side results of the step are still fetchable (pk_gjk_epa_batch would have taken them)
        keys2, contacts2 = ctx.pairs(), ctx.contacts()
        assert np.array_equal(keys, keys2) and np.array_equal(contacts.view(np.uint8), contacts2.view(np.uint8))
        n = len(pa)
        d_a, d_b = ctx.device_alloc(4 * n), ctx.device_alloc(4 * n)
        d_out, d_sep = ctx.device_alloc(64 * n), ctx.device_alloc(n)
        ctx.h2d(d_a, pa)
        ctx.h2d(d_b, pb)
        ms = ctx.gjk_distance_batch_device(d_a, d_b, n, d_out, d_sep)
        rec2 = np.zeros_like(rec)
        sep2 = np.zeros_like(sep)
        ctx.d2h(rec2, d_out)
        ctx.d2h(sep2, d_sep)
        assert ms > 0.0 and np.array_equal(sep, sep2) and np.array_equal(rec.view(np.uint8), rec2.view(np.uint8))
        for p in (d_a, d_b, d_out, d_sep):
            ctx.device_free(p)
    finally:
        ctx.close()


def test_distance_rejects_bad_indices_and_handles_empty_lists():
    import physkit_b200 as pk

    sc, pa, pb = random_pairs_scene(8, 1)
    ctx = make_context(sc, 16)
    try:
        sep, rec = ctx.gjk_distance_batch([], [])
        assert len(sep) == 0 and len(rec) == 0
        with pytest.raises(pk.PkError):
            ctx.gjk_distance_batch([0], [sc.n])
    finally:
        ctx.close()
