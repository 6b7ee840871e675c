// The exchange step of a world that spans GPUs (include/pk_collide.h, pk_comm_*), driven from C++ the way a host engine
// would: N contexts on the N devices it finds (one host thread per rank, ranks of ONE communicator), poses uploaded by
// slice and all-gathered, one step, contact records all-gathered — four rounds: the first exchange (block size from
// the counts), a second of about the same size (block size guessed from the first), a much denser scene (the guess is
// outgrown: the records go out a second time) and the first scene again (the guess is far too large).  With one device
// it is a communicator of one rank.
// Checks per round: every rank ends up with the same N blocks; block r is rank r's own contact list; the union of the
// ranks' pair sets has no duplicates.  Exit code: 0 ok, 3 no CUDA device / no NCCL, 1 failure.
#include "pk_collide.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

static uint64_t mix(uint64_t &s)
{
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double u01(uint64_t &s) { return static_cast<double>(mix(s) >> 11) * (1.0 / 9007199254740992.0); }

int main()
{
    int ndev = 0; // devices the library can open (a small probe context each), at most four ranks
    for (; ndev < 4; ++ndev)
    {
        pk_config probe;
        std::memset(&probe, 0, sizeof(probe));
        probe.device = ndev;
        probe.mode = PK_MODE_QUERY;
        probe.max_bodies = 2;
        probe.max_shapes = 1;
        probe.max_pairs = 16;
        probe.num_worlds = 1;
        probe.shard_count = 1;
        pk_ctx *c = nullptr;
        if (pk_create(&probe, &c) != PK_OK) break;
        pk_destroy(c);
    }
    if (ndev < 1) return std::printf("no CUDA device\n"), 3;
    const int N = ndev;
    pk_comm_id id;
    if (pk_comm_get_id(&id) != PK_OK) return std::printf("no NCCL\n"), 3;
    const uint32_t side = 12, n = side * side * side;
    std::vector<double> pos(3 * n), quat(4 * n), disp(3 * n, 0.0);
    std::vector<uint32_t> sid(n);
    std::vector<uint8_t> flags(n, 2);
    uint64_t seed = 42;
    for (uint32_t i = 0; i < n; ++i)
    {
        pos[3 * i] = 0.8 * (i % side) + 0.2 * u01(seed);
        pos[3 * i + 1] = 0.8 * ((i / side) % side) + 0.2 * u01(seed);
        pos[3 * i + 2] = 0.8 * (i / (side * side)) + 0.2 * u01(seed);
        quat[4 * i] = quat[4 * i + 1] = quat[4 * i + 2] = 0.0;
        quat[4 * i + 3] = 1.0;
        sid[i] = i & 1u;
    }
    std::vector<int> status(N, 0);
    constexpr int ROUNDS = 4;
    const double shift[ROUNDS] = {0.05, 0.08, 0.0, 0.05}, scale[ROUNDS] = {1.0, 1.0, 0.75, 1.0};
    std::vector<std::vector<pk_contact>> own_r[ROUNDS], gathered_r[ROUNDS];
    std::vector<std::vector<uint64_t>> counts_r[ROUNDS], keys_r[ROUNDS];
    for (int k = 0; k < ROUNDS; ++k)
    {
        own_r[k].resize(N);
        gathered_r[k].resize(N);
        counts_r[k].resize(N);
        keys_r[k].resize(N);
    }
    auto rank_main = [&](int r)
    {
        auto fail = [&](const char *what, int s)
        {
            std::printf("rank %d: %s failed (%d)\n", r, what, s);
            status[r] = 1;
        };
        pk_config cfg;
        std::memset(&cfg, 0, sizeof(cfg));
        cfg.device = r;
        cfg.mode = PK_MODE_WORLD;
        cfg.max_bodies = n;
        cfg.max_shapes = 4;
        cfg.max_pairs = 200000;
        cfg.num_worlds = 1;
        cfg.shard_rank = static_cast<uint32_t>(r);
        cfg.shard_count = static_cast<uint32_t>(N);
        pk_ctx *ctx = nullptr;
        int s = pk_create(&cfg, &ctx);
        if (s != PK_OK) return fail("pk_create", s);
        uint32_t sph, box;
        const double half[3] = {0.3, 0.35, 0.4};
        pk_shape_sphere(ctx, 0.4, &sph);
        pk_shape_box(ctx, half, &box);
        if ((s = pk_comm_init(ctx, &id, r, N)) != PK_OK) return fail("pk_comm_init", s);
        pk_bodies_resize(ctx, n);
        // shape ids and flags are not per-step data: every rank uploads them once, with the first poses
        if ((s = pk_bodies_upload(ctx, pos.data(), quat.data(), disp.data(), sid.data(), flags.data(), nullptr, 0, n)) != PK_OK)
            return fail("pk_bodies_upload", s);
        pk_step_result res;
        if ((s = pk_collide_resident(ctx, &res)) != PK_OK) return fail("first step", s); // creates the bodies: no pairs yet
        // the steps proper: this rank moves ITS slice of the bodies, the others arrive over NVLink
        uint32_t first = 0, count = 0;
        pk_comm_pose_slice(ctx, &first, &count);
        std::vector<double> moved(3 * count);
        for (int k = 0; k < ROUNDS; ++k)
        {
            auto &own = own_r[k];
            auto &gathered = gathered_r[k];
            auto &counts = counts_r[k];
            auto &keys = keys_r[k];
            for (uint32_t i = 0; i < 3 * count; ++i) moved[i] = pos[3 * first + i] * scale[k] + shift[k];
            if ((s = pk_bodies_update_pose(ctx, moved.data(), nullptr, nullptr, first, count)) != PK_OK) return fail("update_pose", s);
            if ((s = pk_comm_allgather_poses(ctx, PK_POSE_POS)) != PK_OK) return fail("allgather_poses", s);
            if ((s = pk_collide(ctx, &res)) != PK_OK) return fail("pk_collide", s);
            const pk_contact *recs = nullptr;
            uint64_t nrec = 0;
            pk_contacts(ctx, &recs, &nrec);
            own[r].assign(recs, recs + nrec);
            const uint64_t *pk = nullptr;
            uint64_t npk = 0;
            pk_pairs(ctx, &pk, &npk);
            keys[r].assign(pk, pk + npk);
            pk_gathered_contacts g;
            if ((s = pk_comm_allgather_contacts(ctx, &g)) != PK_OK) return fail("allgather_contacts", s);
            counts[r].assign(g.counts, g.counts + g.num_ranks);
            gathered[r].resize(g.total);
            uint64_t off = 0;
            for (uint32_t q = 0; q < g.num_ranks; ++q)
            {
                if (g.counts[q] > g.stride_records) return fail("a block shorter than its count", 1);
                if (g.counts[q])
                    pk_memcpy_d2h(ctx, gathered[r].data() + off, static_cast<const pk_contact *>(g.d_records) + q * g.stride_records,
                                  g.counts[q] * sizeof(pk_contact));
                off += g.counts[q];
            }
        }
        pk_destroy(ctx);
    };
    std::vector<std::thread> th;
    for (int r = 0; r < N; ++r) th.emplace_back(rank_main, r);
    for (auto &t : th) t.join();
    for (int r = 0; r < N; ++r)
        if (status[r]) return 1;
    uint64_t totals[ROUNDS];
    for (int k = 0; k < ROUNDS; ++k)
    {
        auto &own = own_r[k];
        auto &gathered = gathered_r[k];
        auto &counts = counts_r[k];
        auto &keys = keys_r[k];
        uint64_t total = 0;
        std::vector<uint64_t> all_keys;
        for (int r = 0; r < N; ++r)
        {
            total += own[r].size();
            all_keys.insert(all_keys.end(), keys[r].begin(), keys[r].end());
            if (counts[r] != counts[0]) return std::printf("FAIL: ranks disagree about the counts\n"), 1;
            if (counts[r][r] != own[r].size()) return std::printf("FAIL: count of rank %d\n", r), 1;
        }
        std::sort(all_keys.begin(), all_keys.end());
        if (std::adjacent_find(all_keys.begin(), all_keys.end()) != all_keys.end()) return std::printf("FAIL: a pair on two ranks\n"), 1;
        if (total < 200) return std::printf("FAIL: only %llu contacts\n", static_cast<unsigned long long>(total)), 1;
        for (int r = 0; r < N; ++r)
        {
            if (gathered[r].size() != total)
                return std::printf("FAIL: rank %d gathered %zu of %llu\n", r, gathered[r].size(), static_cast<unsigned long long>(total)), 1;
            uint64_t off = 0;
            for (int q = 0; q < N; ++q)
            {
                if (own[q].size() && std::memcmp(gathered[r].data() + off, own[q].data(), own[q].size() * sizeof(pk_contact)) != 0)
                    return std::printf("FAIL: round %d, rank %d's copy of rank %d's records differs\n", k, r, q), 1;
                off += own[q].size();
            }
        }
        totals[k] = total;
        std::printf("round %d: %zu pairs, %llu contacts gathered on every rank\n", k, all_keys.size(), static_cast<unsigned long long>(total));
    }
    if (totals[2] < 2 * totals[1]) return std::printf("FAIL: the dense round did not outgrow the guessed block size\n"), 1;
    std::printf("comm ok: %d rank(s), %d rounds\n", N, ROUNDS);
    return 0;
}
