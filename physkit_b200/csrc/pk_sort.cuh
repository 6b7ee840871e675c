// pk_sort.cuh — LSD radix sort (8 bits per pass) and exclusive scan, written for this pipeline.
//
// The sort is the "K2" (Morton keys + body ids) and "K6" (pair keys) stage of the broadphase.
// Per pass:  radix_hist_kernel (tile histograms)  →  radix_scan_kernel (one block per digit scans the
// tile counts)  →  radix_scatter_kernel (stable scatter).  Inside a tile the ranking is done per warp
// with match.any, so keys stay in registers and the only shared memory is the 8 KB of per-warp digit
// counters.  Only the byte positions that can be non-zero are sorted (pair keys of N bodies need
// ceil(2·log2 N / 8) passes, not 8: sort_digit packs the two id fields).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "pk_common.cuh"

namespace pk
{

constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS; // 4096 keys per block


// Digit of a pass.  Pair keys are (min id << 32) | max id with ids below 2^b: their 2b significant bits sit in
// two fields.  With lowbits = b the fields are packed next to each other before the digit is taken, so that
// a million-body scene (b = 20) sorts its pairs in five 8-bit passes instead of six.  lowbits = 0: plain key.
__device__ __forceinline__ uint32_t sort_digit(uint64_t key, int shift, int lowbits)
{
    if (lowbits) key = (key & ((1ull << lowbits) - 1ull)) | ((key >> 32) << lowbits);
    return static_cast<uint32_t>(key >> shift) & 0xFFu;
}

// item r of warp w, lane l  ↔  tile offset w·(32·ITEMS) + r·32 + l   (warp-contiguous ⇒ stable)
__device__ __forceinline__ uint64_t sort_index(uint64_t tile_base, int warp, int round, int lane)
{
    return tile_base + static_cast<uint64_t>(warp) * (32 * SORT_ITEMS) + static_cast<uint64_t>(round) * 32 + lane;
}

__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(const uint64_t *__restrict__ keys, uint64_t n_cap, const unsigned long long *__restrict__ n_dev, int shift, int lowbits,
                  uint32_t *__restrict__ tile_hist, uint32_t ntiles)
{
    const uint64_t n = device_count(n_dev, n_cap);
    __shared__ uint32_t hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * SORT_TILE;
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r)
    {
        uint64_t i = base + static_cast<uint64_t>(r) * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&hist[sort_digit(keys[i], shift, lowbits)], 1u);
    }
    __syncthreads();
    tile_hist[static_cast<uint64_t>(threadIdx.x) * ntiles + blockIdx.x] = hist[threadIdx.x];
}

// block-wide exclusive scan of one value per thread (256 threads); returns the exclusive prefix and
// leaves the block total in *total.
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t *warp_sums /*[8]*/, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w)
    {
        uint32_t s = warp_sums[w];
        if (w < warp) woff += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return woff + x - v;
}

// grid = 256 blocks (one per digit): exclusive scan of that digit's tile counts, in place.
__global__ void __launch_bounds__(SORT_THREADS)
radix_scan_kernel(uint32_t *__restrict__ tile_hist, uint32_t ntiles, uint32_t *__restrict__ digit_total)
{
    __shared__ uint32_t warp_sums[SORT_WARPS];
    uint32_t *row = tile_hist + static_cast<uint64_t>(blockIdx.x) * ntiles;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < ntiles; base += SORT_THREADS)
    {
        uint32_t i = base + threadIdx.x;
        uint32_t v = (i < ntiles) ? row[i] : 0u;
        uint32_t tot;
        uint32_t ex = block_exclusive_scan_256(v, warp_sums, &tot);
        if (i < ntiles) row[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) digit_total[blockIdx.x] = carry;
}

template <bool HAS_VALS>
__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                     uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint64_t n_cap,
                     const unsigned long long *__restrict__ n_dev, int shift, int lowbits, const uint32_t *__restrict__ tile_hist,
                     uint32_t ntiles, const uint32_t *__restrict__ digit_total)
{
    const uint64_t n = device_count(n_dev, n_cap);
    __shared__ uint32_t warp_cnt[SORT_WARPS][256];
    __shared__ uint32_t warp_sums[SORT_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int w = 0; w < SORT_WARPS; ++w) warp_cnt[w][threadIdx.x] = 0;
    // global base of every digit = exclusive scan of the digit totals + this tile's offset
    uint32_t tot;
    uint32_t digit_base = block_exclusive_scan_256(digit_total[threadIdx.x], warp_sums, &tot) +
                          tile_hist[static_cast<uint64_t>(threadIdx.x) * ntiles + blockIdx.x];
    __syncthreads();

    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * SORT_TILE;
    uint64_t key[SORT_ITEMS];
    uint32_t rank[SORT_ITEMS];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r)
    {
        uint64_t i = sort_index(base, warp, r, lane);
        bool ok = i < n;
        key[r] = ok ? keys_in[i] : 0xFFFFFFFFFFFFFFFFull;
        uint32_t d = sort_digit(key[r], shift, lowbits);
        // out-of-range lanes take a private pseudo-digit so they match nobody real
        uint32_t md = ok ? d : (256u + lane);
        uint32_t peers = __match_any_sync(0xFFFFFFFFu, md);
        uint32_t before = ok ? warp_cnt[warp][d] : 0u;
        __syncwarp();
        rank[r] = before + __popc(peers & lt_mask);
        if (ok && (peers & lt_mask) == 0) warp_cnt[warp][d] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // prefix over warps for digit = threadIdx.x
    {
        uint32_t run = digit_base;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w)
        {
            uint32_t c = warp_cnt[w][threadIdx.x];
            warp_cnt[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r)
    {
        uint64_t i = sort_index(base, warp, r, lane);
        if (i < n)
        {
            uint32_t d = sort_digit(key[r], shift, lowbits);
            uint32_t pos = warp_cnt[warp][d] + rank[r];
            keys_out[pos] = key[r];
            if (HAS_VALS) vals_out[pos] = vals_in[i];
        }
    }
}

// Up to one tile of keys (4096): all passes in ONE launch of one block.  A world of a thousand bodies — the size of the
// reference's own demos — is bound by launch latency, and its body sort was 12 launches (BASELINE C1: 0.076 of the
// 0.37 ms of a step).  Same ranking as radix_scatter_kernel (per warp with match.any, warp-contiguous items: stable),
// the digit bases come from the block's own counts instead of the tile histograms; the passes ping-pong between the
// two buffers like the launches of radix_sort do, a __syncthreads between them (one block: its global writes are
// visible to its own threads after the barrier).  shifts: one byte per pass, lowest pass first.
template <bool HAS_VALS>
__global__ void __launch_bounds__(SORT_THREADS)
radix_sort_tile_kernel(uint64_t *keys0, uint32_t *vals0, uint64_t *keys1, uint32_t *vals1,
                       uint64_t n_cap, const unsigned long long *__restrict__ n_dev, uint64_t shifts, int npasses, int lowbits)
{
    const uint64_t n = device_count(n_dev, n_cap);
    __shared__ uint32_t warp_cnt[SORT_WARPS][256];
    __shared__ uint32_t warp_sums[SORT_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int pass = 0; pass < npasses; ++pass)
    {
        const int shift = static_cast<int>((shifts >> (8 * pass)) & 0xFFull);
        const uint64_t *keys_in = (pass & 1) ? keys1 : keys0;
        const uint32_t *vals_in = (pass & 1) ? vals1 : vals0;
        uint64_t *keys_out = (pass & 1) ? keys0 : keys1;
        uint32_t *vals_out = (pass & 1) ? vals0 : vals1;
        for (int w = 0; w < SORT_WARPS; ++w) warp_cnt[w][threadIdx.x] = 0;
        __syncthreads();
        uint64_t key[SORT_ITEMS];
        uint32_t val[SORT_ITEMS];
        uint32_t rank[SORT_ITEMS];
#pragma unroll
        for (int r = 0; r < SORT_ITEMS; ++r)
        {
            const uint64_t i = sort_index(0, warp, r, lane);
            const bool ok = i < n;
            key[r] = ok ? keys_in[i] : 0xFFFFFFFFFFFFFFFFull;
            val[r] = (HAS_VALS && ok) ? vals_in[i] : 0u;
            const uint32_t d = sort_digit(key[r], shift, lowbits);
            const uint32_t md = ok ? d : (256u + lane); // out-of-range lanes match nobody real
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, md);
            const uint32_t before = ok ? warp_cnt[warp][d] : 0u;
            __syncwarp();
            rank[r] = before + __popc(peers & lt_mask);
            if (ok && (peers & lt_mask) == 0) warp_cnt[warp][d] = before + __popc(peers);
            __syncwarp();
        }
        __syncthreads(); // every key of this pass has been read (the other buffer is about to be overwritten) and counted
        {
            uint32_t c = 0;
#pragma unroll
            for (int w = 0; w < SORT_WARPS; ++w) c += warp_cnt[w][threadIdx.x];
            uint32_t tot;
            uint32_t run = block_exclusive_scan_256(c, warp_sums, &tot); // first position of digit threadIdx.x
#pragma unroll
            for (int w = 0; w < SORT_WARPS; ++w)
            {
                const uint32_t cw = warp_cnt[w][threadIdx.x];
                warp_cnt[w][threadIdx.x] = run;
                run += cw;
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < SORT_ITEMS; ++r)
        {
            const uint64_t i = sort_index(0, warp, r, lane);
            if (i < n)
            {
                const uint32_t pos = warp_cnt[warp][sort_digit(key[r], shift, lowbits)] + rank[r];
                keys_out[pos] = key[r];
                if (HAS_VALS) vals_out[pos] = val[r];
            }
        }
        __syncthreads(); // the pass is complete in global memory before the next one reads it
    }
}

// The hist kernel must count with the same tile partition as the scatter kernel; both cover
// [tile·4096, (tile+1)·4096), only the thread↔item mapping differs, which does not matter for counts.

// ---------------------------------------------------------------------------------------------
// Exclusive scan of u8 flags → u32 (contact slot of every pair).  Three small kernels.
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_TILE = 4096;

__global__ void __launch_bounds__(256)
flag_tile_sum_kernel(const uint8_t *__restrict__ flags, uint64_t n_cap, const unsigned long long *__restrict__ n_dev,
                     uint32_t *__restrict__ tile_sum)
{
    const uint64_t n = device_count(n_dev, n_cap);
    __shared__ uint32_t warp_sums[8];
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * SCAN_TILE + threadIdx.x * 16ull;
    uint32_t c = 0;
    if (base + 16 <= n)
    {
        uint4 v = *reinterpret_cast<const uint4 *>(flags + base);
        c = __popc(v.x & 0x01010101u) + __popc(v.y & 0x01010101u) + __popc(v.z & 0x01010101u) + __popc(v.w & 0x01010101u);
    }
    else
        for (uint64_t i = base; i < n && i < base + 16; ++i) c += flags[i] & 1u;
    uint32_t tot;
    block_exclusive_scan_256(c, warp_sums, &tot);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}

// single block: exclusive scan of the tile sums, total → *total_out
__global__ void __launch_bounds__(256)
tile_sum_scan_kernel(uint32_t *__restrict__ tile_sum, uint32_t ntiles, unsigned long long *__restrict__ total_out)
{
    __shared__ uint32_t warp_sums[8];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < ntiles; base += 256)
    {
        uint32_t i = base + threadIdx.x;
        uint32_t v = (i < ntiles) ? tile_sum[i] : 0u;
        uint32_t tot;
        uint32_t ex = block_exclusive_scan_256(v, warp_sums, &tot);
        if (i < ntiles) tile_sum[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(256)
flag_scan_apply_kernel(const uint8_t *__restrict__ flags, uint64_t n_cap, const unsigned long long *__restrict__ n_dev,
                       const uint32_t *__restrict__ tile_sum, uint32_t *__restrict__ out_index)
{
    const uint64_t n = device_count(n_dev, n_cap);
    __shared__ uint32_t warp_sums[8];
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * SCAN_TILE + threadIdx.x * 16ull;
    uint8_t f[16];
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j)
    {
        uint64_t i = base + j;
        f[j] = (i < n) ? (flags[i] & 1u) : 0u;
        c += f[j];
    }
    uint32_t tot;
    uint32_t ex = block_exclusive_scan_256(c, warp_sums, &tot) + tile_sum[blockIdx.x];
#pragma unroll
    for (int j = 0; j < 16; ++j)
    {
        uint64_t i = base + j;
        if (i < n) out_index[i] = ex;
        ex += f[j];
    }
}

// Stream compaction of 88-byte records by a validity byte (only run when EPA returned nullopt for
// some GJK hit, which is rare): dst[rank(valid)] = src.
__global__ void __launch_bounds__(256)
compact_records_kernel(const uint8_t *__restrict__ valid, const uint32_t *__restrict__ index, uint64_t n,
                       const unsigned char *__restrict__ src, unsigned char *__restrict__ dst, int rec_bytes)
{
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n || !valid[i]) return;
    const uint64_t *s = reinterpret_cast<const uint64_t *>(src + i * rec_bytes);
    uint64_t *d = reinterpret_cast<uint64_t *>(dst + static_cast<uint64_t>(index[i]) * rec_bytes);
    for (int k = 0; k < rec_bytes / 8; ++k) d[k] = s[k];
}

} // namespace pk
