// broad_emul.cpp — TEST INFRASTRUCTURE: the broadphase kernels (physkit_b200/csrc/pk_broadphase.cuh, pk_sort.cuh) compiled
// by g++ through tests/cpp/simt_host.h and launched in the order pk_collide_resident (pk_api.cu) launches them on the
// device: scene reset → bounds + fat rule → Morton keys → body sort → leaves → hierarchy + refit → ropes → self-overlap
// traversal into pair rows (or the list) → rows written out sorted (or the radix sort of the list).
// tests/test_broad_emul.py compares stored boxes, moved counts and the sorted pair set with the oracle's faithful
// incremental dynamic_bvh every step.  Not linked into, nor reachable from, the product library.
#include "simt_host.h"

#include "../../physkit_b200/csrc/pk_broadphase.cuh"

using namespace pk;

namespace
{
struct BroadEmu
{
    uint32_t cap = 0;
    int32_t epoch = 0;
    std::vector<double> stored;
    std::vector<int32_t> last_move, create;
    std::vector<uint8_t> alive;
    std::vector<uint64_t> pairs; // result of the last step, ascending
    unsigned long long moved = 0, row_overflow = 0;
};
inline uint32_t div_up(uint64_t a, uint64_t b) { return static_cast<uint32_t>((a + b - 1) / b); }
inline int bits_for(uint64_t n)
{
    int b = 1;
    while (b < 64 && (1ull << b) < n) ++b;
    return b;
}
} // namespace

extern "C" void *emu_broad_create(uint32_t max_bodies)
{
    auto *e = new BroadEmu;
    e->cap = max_bodies;
    e->stored.assign(6ull * max_bodies, 0.0);
    e->last_move.assign(max_bodies, -1);
    e->create.assign(max_bodies, 0);
    e->alive.assign(max_bodies, 0);
    return e;
}
extern "C" void emu_broad_destroy(void *h) { delete static_cast<BroadEmu *>(h); }

// One step.  world_id may be null (one world).  rows != 0: pairs into per-body rows; else the list + radix sort.
// shard_rank / shard_count: the slice of sorted leaves this "rank" traverses.  Returns the number of pairs, or -1 when a
// row overflowed (the library then repeats the step in list form).
extern "C" long long emu_broad_step(void *h, const ShapeRec *shapes, const double *pos, const double *quat, const double *disp,
                                    const uint32_t *shape_id, const uint8_t *flags, const uint32_t *world_id, uint32_t num_worlds, uint32_t n,
                                    int mode_query, int rows, uint32_t shard_rank, uint32_t shard_count)
{
    BroadEmu &e = *static_cast<BroadEmu *>(h);
    if (n > e.cap) return -2;
    BodyState st{e.stored.data(), e.last_move.data(), e.create.data(), e.alive.data()};
    uint32_t scene[6];
    unsigned long long counters[4] = {0, 0, 0, 0}; // [0] pairs, [1] moved, [2] row overflow
    simt::launch(1, 64, [&]() { scene_reset_kernel(scene, counters, 4); });
    if (n)
        simt::launch(div_up(n, 256), 256,
                     [&]() { bounds_fat_kernel(shapes, pos, quat, disp, shape_id, flags, n, mode_query, e.epoch, st, scene, counters + 1); });
    uint32_t m = 0;
    for (uint32_t i = 0; i < n; ++i) m += (flags[i] & FLAG_ALIVE) ? 1u : 0u;
    e.pairs.clear();
    e.moved = counters[1];
    e.row_overflow = 0;
    e.epoch += 1;
    if (m < 2) return 0;
    WorldTiling wt;
    wt.num_worlds = num_worlds;
    wt.grid = 1;
    while (static_cast<uint64_t>(wt.grid) * wt.grid * wt.grid < wt.num_worlds) ++wt.grid;
    const uint32_t *d_world = num_worlds > 1 ? world_id : nullptr;

    std::vector<uint64_t> bkeys[2] = {std::vector<uint64_t>(n + 1), std::vector<uint64_t>(n + 1)};
    std::vector<uint32_t> bvals[2] = {std::vector<uint32_t>(n + 1), std::vector<uint32_t>(n + 1)};
    simt::launch(div_up(n, 256), 256, [&]() { morton_kernel(e.stored.data(), e.alive.data(), d_world, n, scene, wt, bkeys[0].data(), bvals[0].data()); });
    int bres = 0;
    {
        const int shifts[4] = {0, 8, 16, 24};
        const uint32_t ntiles = div_up(n, SORT_TILE);
        if (ntiles <= 1)
        {
            uint64_t packed = 0;
            for (int k = 0; k < 4; ++k) packed |= static_cast<uint64_t>(shifts[k]) << (8 * k);
            simt::launch(1, SORT_THREADS, [&]()
                         { radix_sort_tile_kernel<true>(bkeys[0].data(), bvals[0].data(), bkeys[1].data(), bvals[1].data(), n, nullptr, packed, 4, 0); });
        }
        else
        {
            std::vector<uint32_t> tile_hist(static_cast<size_t>(ntiles) * 256), digit_total(256);
            for (int p = 0; p < 4; ++p)
            {
                simt::launch(ntiles, SORT_THREADS, [&]() { radix_hist_kernel(bkeys[bres].data(), n, nullptr, shifts[p], 0, tile_hist.data(), ntiles); });
                simt::launch(256, SORT_THREADS, [&]() { radix_scan_kernel(tile_hist.data(), ntiles, digit_total.data()); });
                simt::launch(ntiles, SORT_THREADS,
                             [&]()
                             {
                                 radix_scatter_kernel<true>(bkeys[bres].data(), bvals[bres].data(), bkeys[bres ^ 1].data(), bvals[bres ^ 1].data(), n, nullptr,
                                                            shifts[p], 0, tile_hist.data(), ntiles, digit_total.data());
                             });
                bres ^= 1;
            }
        }
    }
    std::vector<LeafRec> leaves(m);
    std::vector<NodeF> nodes(2ull * m);
    std::vector<int32_t> merge_flag(m);
    std::vector<uint32_t> right(m), range_last(m);
    uint32_t root = 0;
    simt::launch(div_up(m, 256), 256,
                 [&]()
                 {
                     leaf_kernel(bvals[bres].data(), m, e.stored.data(), e.last_move.data(), e.create.data(), d_world, scene, wt, leaves.data(), nodes.data(),
                                 merge_flag.data());
                 });
    simt::launch(div_up(m, 256), 256, [&]() { hierarchy_kernel(bkeys[bres].data(), m, nodes.data(), right.data(), range_last.data(), merge_flag.data(), &root); });
    simt::launch(div_up(2ull * m - 1, 256), 256, [&]() { rope_kernel(m, nodes.data(), right.data(), range_last.data()); });

    const uint32_t p_begin = static_cast<uint32_t>(static_cast<uint64_t>(m) * shard_rank / shard_count);
    const uint32_t p_end = static_cast<uint32_t>(static_cast<uint64_t>(m) * (shard_rank + 1) / shard_count);
    if (p_end <= p_begin) return 0;
    const uint64_t capacity = static_cast<uint64_t>(n) * PAIR_ROW;
    std::vector<uint64_t> pkeys[2] = {std::vector<uint64_t>(capacity + 1), std::vector<uint64_t>(capacity + 1)};
    std::vector<uint32_t> row_count(n, 0u), row_data(static_cast<uint64_t>(n) * PAIR_ROW);
    PairRows pr{row_count.data(), row_data.data(), counters + 2};
    const uint32_t blocks = div_up(p_end - p_begin, OVERLAP_THREADS);
    if (rows)
        simt::launch(blocks, OVERLAP_THREADS,
                     [&]() { overlap_kernel<true>(nodes.data(), leaves.data(), m, p_begin, p_end, mode_query, pkeys[0].data(), capacity, counters + 0, pr); });
    else
        simt::launch(blocks, OVERLAP_THREADS,
                     [&]() { overlap_kernel<false>(nodes.data(), leaves.data(), m, p_begin, p_end, mode_query, pkeys[0].data(), capacity, counters + 0, pr); });
    const uint64_t npairs = counters[0];
    e.row_overflow = counters[2];
    if (npairs > capacity) return -3;
    if (rows && counters[2]) return -1;
    int pair_buf = 0;
    if (npairs && rows)
    {
        const uint32_t nt = div_up(n, ROWS_TILE);
        std::vector<uint32_t> row_tiles(nt + 1);
        simt::launch(nt, ROWS_TILE, [&]() { pair_rows_sum_kernel(row_count.data(), n, row_tiles.data()); });
        simt::launch(1, 256, [&]() { tile_sum_scan_kernel(row_tiles.data(), nt, nullptr); });
        simt::launch(nt, ROWS_TILE, [&]() { pair_rows_emit_kernel(row_count.data(), row_data.data(), n, row_tiles.data(), pkeys[1].data(), capacity); });
        pair_buf = 1;
    }
    else if (npairs)
    {
        const int idbits = bits_for(n);
        std::vector<int> shifts;
        for (int b = 0; b < 2 * idbits; b += 8) shifts.push_back(b);
        const uint32_t ntiles = div_up(npairs, SORT_TILE);
        if (ntiles <= 1)
        {
            uint64_t packed = 0;
            for (size_t k = 0; k < shifts.size(); ++k) packed |= static_cast<uint64_t>(shifts[k]) << (8 * k);
            simt::launch(1, SORT_THREADS, [&]()
                         { radix_sort_tile_kernel<false>(pkeys[0].data(), nullptr, pkeys[1].data(), nullptr, npairs, nullptr, packed, static_cast<int>(shifts.size()), idbits); });
            pair_buf = static_cast<int>(shifts.size() & 1);
        }
        else
        {
            std::vector<uint32_t> tile_hist(static_cast<size_t>(ntiles) * 256), digit_total(256);
            for (int shift : shifts)
            {
                simt::launch(ntiles, SORT_THREADS, [&]() { radix_hist_kernel(pkeys[pair_buf].data(), npairs, nullptr, shift, idbits, tile_hist.data(), ntiles); });
                simt::launch(256, SORT_THREADS, [&]() { radix_scan_kernel(tile_hist.data(), ntiles, digit_total.data()); });
                simt::launch(ntiles, SORT_THREADS,
                             [&]()
                             {
                                 radix_scatter_kernel<false>(pkeys[pair_buf].data(), nullptr, pkeys[pair_buf ^ 1].data(), nullptr, npairs, nullptr, shift, idbits,
                                                             tile_hist.data(), ntiles, digit_total.data());
                             });
                pair_buf ^= 1;
            }
        }
    }
    e.pairs.assign(pkeys[pair_buf].begin(), pkeys[pair_buf].begin() + npairs);
    return static_cast<long long>(npairs);
}

extern "C" void emu_broad_get(void *h, uint64_t *pairs, double *stored6, uint32_t n, unsigned long long *moved)
{
    BroadEmu &e = *static_cast<BroadEmu *>(h);
    if (pairs) std::memcpy(pairs, e.pairs.data(), e.pairs.size() * sizeof(uint64_t));
    if (stored6) std::memcpy(stored6, e.stored.data(), 6ull * n * sizeof(double));
    if (moved) *moved = e.moved;
}
