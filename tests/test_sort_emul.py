"""The library's radix sort and flag scan (physkit_b200/csrc/pk_sort.cuh), their own source run on the host through
tests/emul.py (one OS thread per CUDA thread, warp votes / match.any / block barriers as barriers), against numpy's
stable sort and cumulative sum.  The sort orders the bodies along the Morton curve (K2), the pair keys of the list form
of the broadphase (K6), ray hits and manifold candidates; the scan turns GJK hit flags into contact slots."""
import numpy as np
import pytest

import emul

pytestmark = pytest.mark.skipif(not emul.available(), reason="CUDA headers not installed")


def _expect(keys, vals, mask):
    order = np.argsort(keys & np.uint64(mask), kind="stable")
    return keys[order], vals[order]


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 511, 513, 1001, 4095, 4096])
def test_tile_sort_is_a_stable_sort(n):
    """radix_sort_tile_kernel: all passes of a sort of up to 4096 keys in one launch of one block."""
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << 30, n, dtype=np.uint64)
    keys[:: 3] = keys[0]  # many equal keys: their values must keep their order
    vals = np.arange(n, dtype=np.uint32)
    k, v = emul.radix_sort(keys, vals, [0, 8, 16, 24])
    ek, ev = _expect(keys, vals, (1 << 32) - 1)
    assert np.array_equal(k, ek) and np.array_equal(v, ev)


def test_tile_sort_odd_pass_counts_keys_only_and_device_side_count():
    rng = np.random.default_rng(7)
    keys = rng.integers(0, 1 << 24, 3000, dtype=np.uint64)
    k, v = emul.radix_sort(keys, None, [0, 8, 16])  # three passes: the result ends in the second buffer
    assert v is None and np.array_equal(k, np.sort(keys))
    vals = np.arange(3000, dtype=np.uint32)
    k, v = emul.radix_sort(keys, vals, [0, 8, 16], n_dev=1234)  # the count lives on the device: the rest is not touched
    ek, ev = _expect(keys[:1234], vals[:1234], (1 << 24) - 1)
    assert np.array_equal(k[:1234], ek) and np.array_equal(v[:1234], ev)
    assert np.array_equal(k[1234:], keys[1234:]) and np.array_equal(v[1234:], vals[1234:])


def test_pair_keys_sort_with_packed_id_fields():
    """Pair keys (min id << 32) | max id of a scene of 2^b bodies: lowbits = b packs the two fields, so 2b bits are sorted
    (pk_api.cu: pair_sort_shifts) and the result is the order of the full 64-bit keys."""
    rng = np.random.default_rng(11)
    b = 10
    a = rng.integers(0, 1 << b, 4000, dtype=np.uint64)
    c = rng.integers(0, 1 << b, 4000, dtype=np.uint64)
    keys = (np.minimum(a, c) << np.uint64(32)) | np.maximum(a, c)
    k, _ = emul.radix_sort(keys, None, [0, 8, 16], lowbits=b)
    assert np.array_equal(k, np.sort(keys))


def test_multi_launch_pass_over_several_tiles_and_against_the_tile_sort():
    """radix_hist / radix_scan / radix_scatter: one pass over three tiles is a stable sort by that byte; on one tile it
    gives what the single-launch form gives."""
    rng = np.random.default_rng(3)
    keys = rng.integers(0, 1 << 16, 9000, dtype=np.uint64)
    keys[::5] = 77 << 8
    vals = np.arange(9000, dtype=np.uint32)
    k, v = emul.radix_sort(keys, vals, [8], tile=False)
    order = np.argsort((keys >> np.uint64(8)) & np.uint64(0xFF), kind="stable")
    assert np.array_equal(k, keys[order]) and np.array_equal(v, vals[order])
    k1, v1 = emul.radix_sort(keys[:2500], vals[:2500], [8], tile=False)
    k2, v2 = emul.radix_sort(keys[:2500], vals[:2500], [8], tile=True)
    assert np.array_equal(k1, k2) and np.array_equal(v1, v2)


@pytest.mark.parametrize("n,n_dev", [(1, None), (15, None), (4096, None), (4097, None), (10_000, None), (10_000, 6001)])
def test_flag_scan_gives_the_rank_among_the_hits(n, n_dev):
    rng = np.random.default_rng(n)
    flags = (rng.random(n) < 0.3).astype(np.uint8)
    out, total = emul.flag_scan(flags, n_dev=n_dev)
    m = n if n_dev is None else n_dev
    want = np.concatenate([[0], np.cumsum(flags[:m])[:-1]]).astype(np.uint32)
    assert np.array_equal(out[:m], want) and total == int(flags[:m].sum())
