// sort_emul.cpp — TEST INFRASTRUCTURE: the library's radix sort and flag scan (physkit_b200/csrc/pk_sort.cuh) compiled
// by g++ through tests/cpp/simt_host.h and driven the way radix_sort / run_narrowphase (pk_api.cu) drive them on the
// device.  tests/test_sort_emul.py compares with numpy's stable sort and cumulative sum.  Not linked into, nor
// reachable from, the product library.
#include "simt_host.h"

#include "../../physkit_b200/csrc/pk_sort.cuh"

using namespace pk;

// keys / vals (vals may be null) are sorted in place over the listed byte shifts.  tile != 0: radix_sort_tile_kernel (all
// passes in one launch; n ≤ SORT_TILE); else one radix_hist / radix_scan / radix_scatter launch per pass.  n_dev < n_cap
// exercises the device-side element count (elements beyond it must stay untouched by the sort).
extern "C" int emu_radix_sort(uint64_t *keys, uint32_t *vals, uint64_t n_cap, uint64_t n_dev_value, const int *shifts, int nshifts, int lowbits, int tile)
{
    std::vector<uint64_t> k1(n_cap + 1, 0xDEADBEEFDEADBEEFull);
    std::vector<uint32_t> v1(n_cap + 1, 0xDEADBEEFu);
    uint64_t *kb[2] = {keys, k1.data()};
    uint32_t *vb[2] = {vals, v1.data()};
    unsigned long long n_dev = n_dev_value;
    const unsigned long long *n_dev_ptr = n_dev_value <= n_cap ? &n_dev : nullptr;
    int cur = 0;
    if (tile)
    {
        if (n_cap > static_cast<uint64_t>(SORT_TILE) || nshifts > 8) return -1;
        uint64_t packed = 0;
        for (int k = 0; k < nshifts; ++k) packed |= static_cast<uint64_t>(shifts[k] & 0xFF) << (8 * k);
        if (vals)
            simt::launch(1, SORT_THREADS, [&]() { radix_sort_tile_kernel<true>(kb[0], vb[0], kb[1], vb[1], n_cap, n_dev_ptr, packed, nshifts, lowbits); });
        else
            simt::launch(1, SORT_THREADS, [&]() { radix_sort_tile_kernel<false>(kb[0], nullptr, kb[1], nullptr, n_cap, n_dev_ptr, packed, nshifts, lowbits); });
        cur = nshifts & 1;
    }
    else
    {
        const uint32_t ntiles = static_cast<uint32_t>((n_cap + SORT_TILE - 1) / SORT_TILE);
        std::vector<uint32_t> tile_hist(static_cast<size_t>(ntiles) * 256 + 1), digit_total(256);
        for (int p = 0; p < nshifts; ++p)
        {
            const int shift = shifts[p];
            simt::launch(ntiles, SORT_THREADS, [&]() { radix_hist_kernel(kb[cur], n_cap, n_dev_ptr, shift, lowbits, tile_hist.data(), ntiles); });
            simt::launch(256, SORT_THREADS, [&]() { radix_scan_kernel(tile_hist.data(), ntiles, digit_total.data()); });
            if (vals)
                simt::launch(ntiles, SORT_THREADS, [&]()
                             { radix_scatter_kernel<true>(kb[cur], vb[cur], kb[cur ^ 1], vb[cur ^ 1], n_cap, n_dev_ptr, shift, lowbits, tile_hist.data(), ntiles, digit_total.data()); });
            else
                simt::launch(ntiles, SORT_THREADS, [&]()
                             { radix_scatter_kernel<false>(kb[cur], nullptr, kb[cur ^ 1], nullptr, n_cap, n_dev_ptr, shift, lowbits, tile_hist.data(), ntiles, digit_total.data()); });
            cur ^= 1;
        }
    }
    if (cur == 1)
    {
        const uint64_t n = n_dev_ptr ? n_dev : n_cap;
        std::memcpy(keys, k1.data(), n * sizeof(uint64_t));
        if (vals) std::memcpy(vals, v1.data(), n * sizeof(uint32_t));
    }
    return 0;
}

// contact slot of every pair = exclusive scan of the hit flags (flag_tile_sum / tile_sum_scan / flag_scan_apply)
extern "C" int emu_flag_scan(const uint8_t *flags, uint64_t n_cap, uint64_t n_dev_value, uint32_t *out_index, unsigned long long *total)
{
    unsigned long long n_dev = n_dev_value;
    const unsigned long long *n_dev_ptr = n_dev_value <= n_cap ? &n_dev : nullptr;
    const uint32_t nt = static_cast<uint32_t>((n_cap + SCAN_TILE - 1) / SCAN_TILE);
    std::vector<uint32_t> tiles(nt + 1);
    std::vector<uint8_t> padded(static_cast<size_t>(nt) * SCAN_TILE + 16, 0); // (the kernel reads whole 16-byte groups below n)
    std::memcpy(padded.data(), flags, n_cap);
    simt::launch(nt, 256, [&]() { flag_tile_sum_kernel(padded.data(), n_cap, n_dev_ptr, tiles.data()); });
    simt::launch(1, 256, [&]() { tile_sum_scan_kernel(tiles.data(), nt, total); });
    simt::launch(nt, 256, [&]() { flag_scan_apply_kernel(padded.data(), n_cap, n_dev_ptr, tiles.data(), out_index); });
    return 0;
}
