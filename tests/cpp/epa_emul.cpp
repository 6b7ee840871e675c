// epa_emul.cpp — TEST INFRASTRUCTURE: the narrowphase's EPA kernels (physkit_b200/csrc/pk_epa_coop.cuh,
// pk_narrowphase.cuh) compiled by g++ through tests/cpp/simt_host.h and driven the way
// run_narrowphase (pk_api.cu) drives them on the device: GJK hits → contact slots → cost-class order →
// epa_init_kernel → epa_coop_kernel → epa_kernel (takes what that one handed back).  tests/test_epa_emul.py compares the records with the oracle bit for bit.
// Not linked into, nor reachable from, the product library.
#include "simt_host.h"


#define PK_EC_STATS
namespace pk
{
unsigned long long g_ec_stats[16];
}
#include "../../physkit_b200/csrc/pk_epa_coop.cuh"
#include "../../physkit_b200/csrc/pk_sort.cuh"
#include "../../physkit_b200/csrc/pk_gjk_filter.cuh"
#include "../../physkit_b200/csrc/pk_distance.cuh"

using namespace pk;

extern "C" void emu_stats(unsigned long long *out, int reset)
{
    for (int i = 0; i < 16; ++i)
    {
        out[i] = pk::g_ec_stats[i];
        if (reset) pk::g_ec_stats[i] = 0;
    }
}

// certainly_separated (pk_gjk_filter.cuh) per pair on the host: out[k] = 1 when the FP32 filter would drop pair k
extern "C" int emu_filter(const ShapeRec *shapes, uint64_t nshapes, const double *verts, uint64_t nverts_pool, const double *pos, const double *quat,
                          const uint32_t *shape_id, const uint32_t *pa, const uint32_t *pb, uint64_t n, int iters, uint8_t *out)
{
    (void)nshapes;
    std::vector<float4> vf(nverts_pool + 2);
    for (uint64_t i = 0; i < nverts_pool; ++i)
        vf[i] = make_float4(static_cast<float>(verts[3 * i]), static_cast<float>(verts[3 * i + 1]), static_cast<float>(verts[3 * i + 2]), 0.f);
    BodyArrays ba;
    ba.shapes = shapes;
    ba.verts = verts;
    ba.verts_f = vf.data();
    ba.pos = pos;
    ba.quat = quat;
    ba.shape_id = shape_id;
    for (uint64_t k = 0; k < n; ++k) out[k] = certainly_separated(load_shape(ba, pa[k]), load_shape(ba, pb[k]), iters) ? 1 : 0;
    return 0;
}

// gjk_distance_pair (pk_distance.cuh) per pair on the host, as gjk_distance_kernel<false> runs it (support<false>: the
// plain FP64 hull scan — the float-prefiltered scan of the <true> instance returns the same vertex, pk_common.cuh)
extern "C" int emu_distance(const ShapeRec *shapes, uint64_t nshapes, const double *verts, uint64_t nverts_pool, const double *pos, const double *quat,
                            const uint32_t *shape_id, const uint32_t *pa, const uint32_t *pb, uint64_t n, DistanceRec *out, uint8_t *separated)
{
    (void)nshapes;
    std::vector<float4> vf(nverts_pool + 2);
    for (uint64_t i = 0; i < nverts_pool; ++i)
        vf[i] = make_float4(static_cast<float>(verts[3 * i]), static_cast<float>(verts[3 * i + 1]), static_cast<float>(verts[3 * i + 2]), 0.f);
    BodyArrays ba;
    ba.shapes = shapes;
    ba.verts = verts;
    ba.verts_f = vf.data();
    ba.pos = pos;
    ba.quat = quat;
    ba.shape_id = shape_id;
    for (uint64_t k = 0; k < n; ++k)
    {
        DistanceRec r;
        r.key = (static_cast<uint64_t>(pa[k]) << 32) | pb[k];
        d3 a, b;
        separated[k] = gjk_distance_pair<false>(load_shape(ba, pa[k]), load_shape(ba, pb[k]), r.distance, a, b) ? 1 : 0;
        r.point_a[0] = a.x; r.point_a[1] = a.y; r.point_a[2] = a.z;
        r.point_b[0] = b.x; r.point_b[1] = b.y; r.point_b[2] = b.z;
        out[k] = r;
    }
    return 0;
}

// arrival = 3: no host-side stand-in for the front of the stage at all — gjk_filter_kernel / gjk_prefilter_kernel,
// gjk_kernel, the flag scan and epa_order_kernel are launched as run_narrowphase launches them (the hits then take their
// simplex slots in whatever order the OS schedules the threads: the records must not depend on it).
extern "C" int emu_gjk_epa(const ShapeRec *shapes, uint64_t nshapes, const double *verts, uint64_t nverts_pool, const double *pos, const double *quat,
                           const uint32_t *shape_id, const uint32_t *pa, const uint32_t *pb, uint64_t n, uint64_t capacity,
                           ContactRec *out, uint8_t *hit_out, int mirror, int arrival, int nblocks, uint64_t *stats /*[8]*/)
{
    std::vector<float4> vf(nverts_pool + 2);
    for (uint64_t i = 0; i < nverts_pool; ++i)
        vf[i] = make_float4(static_cast<float>(verts[3 * i]), static_cast<float>(verts[3 * i + 1]), static_cast<float>(verts[3 * i + 2]), 0.f);
    bool big_hulls = false;
    for (uint64_t k = 0; k < nshapes; ++k) big_hulls = big_hulls || (shapes[k].kind == KIND_HULL && shapes[k].nverts > HULL_PREFILTER_MIN);
    BodyArrays ba;
    ba.shapes = shapes;
    ba.verts = verts;
    ba.verts_f = vf.data();
    ba.pos = pos;
    ba.quat = quat;
    ba.shape_id = shape_id;

    // gjk_kernel (pk_narrowphase.cuh:311-360) per pair.  On the device the hits take their simplex slots in arrival
    // order; `arrival` picks pair order (0), reversed (1) or a fixed pseudo-random permutation (2)
    std::vector<uint8_t> hit(n + SCAN_TILE + 16, 0); // (the flag scan reads whole 16-byte groups)
    std::vector<SimplexRec> simplices(capacity + 1);
    unsigned long long counters[8] = {0, 0, 0, 0, 0, 0, 0, 0}; // [0] hits, [4] valid, [5] dropped
    unsigned long long class_counts[EPA_CLASSES] = {};
    std::vector<uint64_t> seq(n);
    for (uint64_t k = 0; k < n; ++k) seq[k] = arrival == 1 ? n - 1 - k : k;
    if (arrival == 2)
    {
        uint64_t x = 0x9E3779B97F4A7C15ull;
        for (uint64_t k = n; k > 1; --k)
        {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            std::swap(seq[k - 1], seq[x % k]);
        }
    }
    const bool kernels = arrival == 3;
    std::vector<uint32_t> out_index(n + 1, 0);
    std::vector<uint32_t> order(capacity + 1);
    if (kernels && n)
    {
        std::vector<uint32_t> work(4 * (n + 1));
        unsigned long long work_count[4] = {0, 0, 0, 0};
        const unsigned pair_blocks = static_cast<unsigned>((n + 127) / 128);
        const unsigned gjk_blocks = static_cast<unsigned>((n + PK_GJK_THREADS - 1) / PK_GJK_THREADS);
        if (!big_hulls)
        {
            simt::launch(pair_blocks, 128, [&]() { gjk_filter_kernel(ba, nullptr, pa, pb, n, nullptr, hit.data(), work.data(), n + 1, work_count, PK_GJK_FILTER_ITERS); });
            simt::launch(gjk_blocks, PK_GJK_THREADS,
                         [&]()
                         {
                             gjk_kernel<false, false>(ba, nullptr, pa, pb, work.data(), n + 1, work_count, hit.data(), simplices.data(), &counters[0], capacity,
                                                      class_counts, nullptr);
                         });
        }
        else
        {
            std::vector<GjkCarry> carry(n + 1);
            simt::launch(pair_blocks, 128,
                         [&]() { gjk_prefilter_kernel<true, true>(ba, nullptr, pa, pb, n, nullptr, hit.data(), work.data(), n + 1, work_count, carry.data()); });
            simt::launch(gjk_blocks, PK_GJK_THREADS,
                         [&]()
                         {
                             gjk_kernel<true, true>(ba, nullptr, pa, pb, work.data(), n + 1, work_count, hit.data(), simplices.data(), &counters[0], capacity,
                                                    class_counts, carry.data());
                         });
        }
        const unsigned nt = static_cast<unsigned>((n + SCAN_TILE - 1) / SCAN_TILE);
        std::vector<uint32_t> tiles(nt + 1);
        unsigned long long scan_total = 0;
        simt::launch(nt, 256, [&]() { flag_tile_sum_kernel(hit.data(), n, nullptr, tiles.data()); });
        simt::launch(1, 256, [&]() { tile_sum_scan_kernel(tiles.data(), nt, &scan_total); });
        simt::launch(nt, 256, [&]() { flag_scan_apply_kernel(hit.data(), n, nullptr, tiles.data(), out_index.data()); });
        if (scan_total != counters[0]) return -3; // the scan counts the hits gjk_kernel counted
        unsigned long long class_fill[EPA_CLASSES] = {};
        simt::launch(4, 256, [&]() { epa_order_kernel(simplices.data(), &counters[0], capacity, class_counts, class_fill, order.data()); });
    }
    for (uint64_t q = 0; q < n && !kernels; ++q)
    {
        const uint64_t k = seq[q];
        ShapeView A = load_shape(ba, pa[k]);
        ShapeView B = load_shape(ba, pb[k]);
        // gjk_prefilter_kernel + gjk_kernel: the first two supports, the separation test, then gjk_resume
        Simplex s;
        bool h;
        {
            const SupportPt s0 = minkowski_support(A, B, d3{1.0, 0.0, 0.0});
            const d3 p0 = P(s0);
            if (sqnorm(p0) < 1e-12)
            {
                s.pt[0] = s0; // the one-point simplex of collision.cpp:174
                s.n = 1;
                h = true;
            }
            else
            {
                const d3 dir = -normalized(p0);
                const SupportPt s1 = minkowski_support(A, B, dir);
                if (dot(P(s1), dir) <= 0.0)
                    h = false;
                else
                {
                    s.pt[0] = s0;
                    s.pt[1] = s1;
                    s.n = 2;
                    h = gjk_resume(A, B, s);
                }
            }
        }
        hit[k] = h ? 1 : 0;
        if (!h) continue;
        const unsigned long long slot = counters[0]++;
        if (slot >= capacity) continue;
        SimplexRec &r = simplices[slot];
        std::memset(&r, 0, sizeof(r));
        for (int i = 0; i < s.n; ++i)
        {
            r.v[i][0] = s.pt[i].pa.x; r.v[i][1] = s.pt[i].pa.y; r.v[i][2] = s.pt[i].pa.z;
            r.v[i][3] = s.pt[i].pb.x; r.v[i][4] = s.pt[i].pb.y; r.v[i][5] = s.pt[i].pb.z;
        }
        const uint32_t cls = epa_cost_class(A, B);
        class_counts[cls]++;
        r.n = static_cast<uint32_t>(s.n) | (cls << 8);
        r.pair = static_cast<uint32_t>(k);
    }
    const uint64_t nhits = std::min<uint64_t>(counters[0], capacity);
    // contact slot = rank of the pair among the hits; order[] = hit slots grouped by class, heaviest first
    if (!kernels)
    {
        uint32_t run = 0;
        for (uint64_t k = 0; k < n; ++k)
        {
            out_index[k] = run;
            run += hit[k];
        }
    }
    if (!kernels)
    {
        uint64_t fill[EPA_CLASSES]; // heaviest class first, as epa_order_kernel
        uint64_t run = 0;
        for (int c = static_cast<int>(EPA_CLASSES) - 1; c >= 0; --c)
        {
            fill[c] = run;
            run += class_counts[c];
        }
        for (uint64_t s = 0; s < nhits; ++s)
        {
            const uint32_t cls = (simplices[s].n >> 8) & 0xFu;
            if (fill[cls] < nhits) order[fill[cls]] = static_cast<uint32_t>(s);
            fill[cls]++;
        }
    }
    std::vector<EpaInit> init(nhits + 1);
    std::vector<ContactRec> contacts(capacity + 1), host_mirror(capacity + 1);
    std::vector<uint8_t> valid(capacity + 16, 0);
    std::vector<uint32_t> fb2(nhits + 1, 0);
    unsigned long long cursors[4] = {0, 0, 0, 0}, fbc[3] = {0, 0, 0};
    unsigned long long *hit_count = &counters[0];
    const unsigned block = ES_THREADS;
    std::vector<unsigned char> spill(static_cast<size_t>(block) * nblocks * es_slab_bytes());
    std::vector<unsigned char> slabs(static_cast<size_t>(EPA_THREADS) * EPA_SLAB_BYTES);
    ContactRec *mir = mirror ? host_mirror.data() : nullptr;

    if (nhits)
    {
        simt::launch((nhits + 127) / 128, 128,
                     [&]() { epa_init_kernel(simplices.data(), hit_count, capacity, nullptr, pa, pb, out_index.data(), init.data()); });
        EcParams ep;
        ep.bodies = ba;
        ep.simplices = simplices.data();
        ep.hit_count_ptr = hit_count;
        ep.hit_capacity = capacity;
        ep.order = order.data();
        ep.contacts = contacts.data();
        ep.valid = valid.data();
        ep.slabs = spill.data();
        ep.cursor = &cursors[0];
        ep.counters = &counters[4];
        ep.fallback = fb2.data();
        ep.fallback_count = &fbc[1];
        ep.restart_count = &fbc[0];
        ep.init = init.data();
        ep.contacts_host = mir;
        // blocks run one after the other (see simt_host.h): the first takes the whole list, later blocks find it drained
        // the instance the library would pick: HULLS when a hull above HULL_PREFILTER_MIN vertices is registered
        if (mirror && big_hulls)
            simt::launch(nblocks, block, [&]() { epa_coop_kernel<true, true>(ep); });
        else if (mirror)
            simt::launch(nblocks, block, [&]() { epa_coop_kernel<true, false>(ep); });
        else if (big_hulls)
            simt::launch(nblocks, block, [&]() { epa_coop_kernel<false, true>(ep); });
        else
            simt::launch(nblocks, block, [&]() { epa_coop_kernel<false, false>(ep); });
        simt::launch(1, EPA_THREADS,
                     [&]()
                     {
                         epa_kernel(ba, nullptr, pa, pb, simplices.data(), &fbc[1], capacity, out_index.data(), fb2.data(), contacts.data(),
                                    valid.data(), slabs.data(), &cursors[3], &counters[4], mir);
                     });
    }
    // expand_contacts_kernel
    for (uint64_t k = 0; k < n; ++k)
    {
        bool h = hit[k] != 0;
        const uint32_t slot = out_index[k];
        if (h) h = slot < capacity && valid[slot] != 0;
        ContactRec r;
        if (h)
        {
            r = contacts[slot];
            if (mirror && std::memcmp(&r, &host_mirror[slot], sizeof(r)) != 0) return -2; // the pinned copy must be the same record
        }
        else
        {
            std::memset(&r, 0, sizeof(r));
            r.key = (static_cast<uint64_t>(pa[k]) << 32) | pb[k];
        }
        out[k] = r;
        hit_out[k] = h ? 1 : 0;
    }
    if (stats)
    {
        stats[0] = counters[0]; // GJK hits
        stats[1] = fbc[0];      // SCAN pairs started again in HEAP mode
        stats[2] = fbc[1];      // handed to epa_kernel
        stats[3] = counters[4]; // valid contacts
        stats[4] = counters[5]; // dropped: no room for the contact
        stats[5] = class_counts[0]; // no smooth shape
        stats[6] = class_counts[1] + class_counts[2] + class_counts[3] + class_counts[4]; // one
        stats[7] = class_counts[5] + class_counts[6] + class_counts[7] + class_counts[8]; // two
    }
    return 0;
}
