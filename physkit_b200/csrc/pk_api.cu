// pk_api.cu — context, memory and the C ABI of include/pk_collide.h.
//
// One ctx = one device, one stream.  All hot-path state lives in HBM for the lifetime of the ctx
// (SoA body arrays, shape table, LBVH, pair keys, simplices, EPA slabs, contacts); a step touches
// the host only to read two counters.  There is no CPU implementation of any stage in this file:
// if no CUDA device is usable pk_create fails and every other entry point needs a ctx.
#include "../../include/pk_collide.h"

#include "pk_broadphase.cuh"
#include "pk_comm.cuh"
#include "pk_common.cuh"
#include "pk_narrowphase.cuh"
#include "pk_epa_coop.cuh"
#include "pk_gjk_filter.cuh"
#include "pk_distance.cuh"
#include "pk_manifold.cuh"
#include "pk_dynamics.cuh"
#include "pk_ray.cuh"
#include "pk_solver.cuh"
#include "pk_sort.cuh"

#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

using namespace pk;

// EPA implementation: epa_coop_kernel (pk_epa_coop.cuh) + epa_kernel (pk_narrowphase.cuh) for what it hands back
namespace
{

enum Stage
{
    ST_BOUNDS = 0,
    ST_MORTON,
    ST_BODY_SORT,
    ST_LEAVES,
    ST_HIERARCHY,
    ST_OVERLAP,
    ST_PAIR_SORT,
    ST_GJK,
    ST_SCAN,
    ST_EPA,
    ST_COMPACT,
    ST_FETCH,
    ST_COUNT
};
static_assert(ST_COUNT == PK_NUM_STAGES, "stage table out of sync with pk_collide.h");
const char *const kStageNames[ST_COUNT] = {"bounds_fat", "morton",    "body_sort", "leaves",  "hierarchy_ropes", "overlap",
                                           "pair_sort",  "gjk",       "hit_scan",  "epa",     "compact",         "fetch_d2h"};

// counters in device memory (unsigned long long each)
enum Counter
{
    C_PAIRS = 0,
    C_MOVED,
    C_HITS,
    C_EPA_CURSOR,
    C_VALID,
    C_EPA_OVERFLOW,
    C_SCAN_TOTAL,
    C_GJK_CURSOR,
    C_EPA_FALLBACK, // [0] SCAN pairs started again in HEAP mode, [1] pairs epa_coop_kernel handed to epa_kernel
    C_EPA_SCAN_CURSOR = C_EPA_FALLBACK + 3, // work cursor of epa_coop_kernel
    C_EPA_FB_CURSOR = C_EPA_SCAN_CURSOR + 3, // work cursor of epa_kernel on the second fallback list
    C_EPA_REASONS = C_EPA_FALLBACK + 10, // [6] debug builds (PK_ES_REASONS)
    C_GJK_CLASS = 32,   // [4] GJK prefilter survivors per shape-kind class
    C_CLASS_COUNT = 40, // [EPA_CLASSES] GJK hits per EPA cost class
    C_CLASS_FILL = 50,  // [EPA_CLASSES]
    C_ROW_OVERFLOW = 59, // pairs that found the row of their lower id full (overlap_kernel<ROWS>)
    C_COUNT = 60
};

} // namespace

struct pk_ctx
{
    pk_config cfg{};
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr; // pk_collide: pair keys go to the host while the narrowphase runs
    cudaEvent_t ev_sorted = nullptr, ev_pairs_copied = nullptr;
    bool want_host_results = false, pairs_in_flight = false;
    // pk_collide / pk_collide_resident size the grids of the pair sort and the narrowphase (and the copy of the pair keys
    // to the host) from the previous step's pair count and let the kernels read the count itself on the device (no host
    // round trip in the middle of the step); a step whose count outgrew the guess is run again with the count read back
    bool exact_prefilter = false; // PK_GJK_EXACT_PREFILTER at pk_create
    int filter_iters = PK_GJK_FILTER_ITERS; // PK_GJK_FILTER_ITERS at pk_create
    uint64_t pair_guess = 0;
    bool pair_guess_valid = false, pairs_read_back = false;
    bool contacts_mirrored = false; // pk_collide: the EPA kernels stored this step's contact records into h_contacts as well
    std::string last_error;
    int sm_count = 0;

    // shapes
    std::vector<ShapeRec> h_shapes;
    std::vector<double> h_verts;
    bool shapes_dirty = false;
    bool has_big_hulls = false; // a hull above HULL_PREFILTER_MIN vertices is registered
    bool no_tile_sort = false;  // PK_NO_TILE_SORT=1: sorts of up to one tile take the multi-launch path too (A/B)
    ShapeRec *d_shapes = nullptr;
    double *d_verts = nullptr;
    float4 *d_verts_f = nullptr;

    // bodies
    uint32_t n_bodies = 0;
    std::vector<uint8_t> h_flags;
    uint32_t n_alive = 0;
    bool alive_dirty = true;
    double *d_pos = nullptr, *d_quat = nullptr, *d_disp = nullptr;
    uint32_t *d_shape_id = nullptr, *d_world = nullptr;
    uint8_t *d_flags = nullptr;
    bool have_world = false;
    BodyState st{};

    // broadphase
    uint32_t *d_scene = nullptr;
    unsigned long long *d_counters = nullptr;
    unsigned long long *h_counters = nullptr; // pinned
    uint64_t *d_bkeys[2] = {nullptr, nullptr};
    uint32_t *d_bvals[2] = {nullptr, nullptr};
    uint32_t *d_tile_hist = nullptr, *d_digit_total = nullptr;
    size_t tile_hist_entries = 0;
    LeafRec *d_leaves = nullptr;
    NodeF *d_nodes = nullptr;
    uint32_t *d_right = nullptr, *d_range_last = nullptr, *d_root = nullptr;
    int32_t *d_merge_flag = nullptr;
    // pair rows (overlap_kernel<ROWS>): partners with a larger id, per body; rows_off: a row overflowed once, LIST form from then on
    uint32_t *d_row_count = nullptr, *d_rows = nullptr, *d_row_tiles = nullptr;
    bool rows_off = false;
    uint64_t *d_pkeys[2] = {nullptr, nullptr};
    uint64_t *d_pairs_sorted = nullptr; // points into d_pkeys

    // narrowphase
    uint64_t max_contacts = 0;
    uint8_t *d_hit = nullptr;
    uint32_t *d_out_index = nullptr, *d_scan_tiles = nullptr;
    SimplexRec *d_simplices = nullptr;
    ContactRec *d_contacts[2] = {nullptr, nullptr};
    ContactRec *d_contacts_final = nullptr;
    uint8_t *d_valid = nullptr;
    uint32_t *d_valid_index = nullptr;
    uint32_t *d_epa_order = nullptr;
    uint32_t *d_gjk_work = nullptr;
    GjkCarry *d_gjk_carry = nullptr; // first two support points of the prefilter survivors, by pair index
    unsigned char *d_slabs = nullptr;
    unsigned char *d_epa_spill = nullptr;
    uint32_t *d_epa_fallback2 = nullptr; // hit slots epa_coop_kernel handed to epa_kernel
    EpaInit *d_epa_init = nullptr;
    uint32_t epa_scan_blocks = 0;
    uint32_t epa_blocks = 0;
    uint32_t gjk_blocks = 0;

    // results
    int32_t epoch = 0;
    bool have_results = false, fetched = false;
    bool device_results = false; // the narrowphase buffers still hold the last step's results (pk_gjk_epa_batch re-uses them)
    uint64_t num_pairs = 0, num_contacts = 0;
    uint64_t *h_pairs = nullptr;
    size_t h_pairs_cap = 0;
    pk_contact *h_contacts = nullptr;
    size_t h_contacts_cap = 0;
    // manifolds (pk_manifolds_enable): sorted array of non-empty manifolds, double-buffered
    uint64_t man_cap = 0, man_count = 0, man_began = 0, man_ended = 0;
    int man_cur = 0;
    int32_t man_epoch = -1; // step the manifolds were last updated for
    ManifoldRec *d_man[2] = {nullptr, nullptr};
    ManifoldRec *d_man_stage = nullptr;
    uint64_t *d_man_ckeys[2] = {nullptr, nullptr}; // candidate keys (radix sort double buffer)
    uint32_t *d_man_csrc[2] = {nullptr, nullptr};
    uint64_t *d_man_began[2] = {nullptr, nullptr}, *d_man_ended[2] = {nullptr, nullptr};
    uint8_t *d_man_consumed = nullptr;
    unsigned long long *d_man_counters = nullptr;
    double *d_man_imp = nullptr;
    pk_manifold *h_man = nullptr;
    uint64_t *h_man_events = nullptr; // began keys, then ended keys
    size_t h_man_cap = 0, h_man_events_cap = 0;
    int man_began_buf = 0, man_ended_buf = 0;
    ContactPointRec *d_points = nullptr; // pk_contact_points: allocated on first use
    pk_contact_point *h_points = nullptr;
    size_t points_cap = 0;

    // integrator (pk_dynamics_enable): velocities, force accumulators, mass properties
    bool dyn_enabled = false;
    DynArrays dyn{};
    cudaEvent_t ev_dyn[2]{};
    float dyn_ms = 0.f;
    double *d_material = nullptr; // [n][2] restitution, friction (core/object.h:90-91)
    // contact rows (pk_contact_rows_setup): allocated on first use, 4 point slots per manifold
    uint64_t sol_cap = 0, sol_count = 0;
    uint8_t *d_sol_valid = nullptr;
    uint32_t *d_sol_index = nullptr, *d_sol_tiles = nullptr, *d_sol_slot = nullptr;
    SolverPoint *d_sol_rows = nullptr;
    unsigned long long *d_sol_total = nullptr;
    pk_solver_point *h_sol_rows = nullptr;
    size_t h_sol_cap = 0;
    cudaEvent_t ev_sol[2]{};
    float sol_ms = 0.f;

    // ray casts (pk_raycast): buffers allocated on first use
    uint32_t tree_m = 0;      // leaves of the tree the last step built (0: none, < 2 bodies alive)
    bool tree_valid = false;  // a step has run since the context was created
    double *d_ray_o = nullptr, *d_ray_d = nullptr, *d_ray_max = nullptr, *d_ray_dist = nullptr;
    uint32_t *d_ray_world = nullptr;
    size_t ray_cap = 0, ray_hit_cap = 0;
    uint64_t *d_ray_keys[2] = {nullptr, nullptr};
    uint32_t *d_ray_src[2] = {nullptr, nullptr};
    RayHitRec *d_ray_hits = nullptr;
    unsigned long long *d_ray_counter = nullptr;
    cudaEvent_t ev_ray[2]{};
    float ray_ms = 0.f;

    cudaEvent_t ev[ST_COUNT + 1]{};
    float stage_ms[ST_COUNT]{};
    uint32_t launches = 0;

    // one world over several processes (pk_comm_*): NCCL communicator on this context's stream
    ncclComm_t comm = nullptr;
    int comm_rank = 0, comm_size = 1;
    unsigned long long *d_comm_counts = nullptr; // [comm_size + 1]: every rank's contact count (+ this rank's, as sent)
    unsigned long long *h_comm_counts = nullptr; // pinned
    ContactRec *d_gather = nullptr;              // room for comm_size blocks of gather_stride records
    uint64_t gather_stride = 0;
    uint64_t gather_guess = 0; // largest per-rank count of the previous exchange (0: none yet)
    cudaEvent_t ev_comm[2]{};
};

namespace
{

// Self-test of pk_div_by_rcp (pk_common.cuh) against IEEE division: every thread draws `per_thread` operand
// pairs from a counter-based generator and counts quotients that differ in any bit.  Families of operands:
// random mantissas; numerators constructed next to an exact multiple of the divisor (quotients next to a
// representable value or next to a rounding boundary); divisors with all-ones / all-zero mantissas.
__device__ __forceinline__ uint64_t st_mix(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void division_selftest_kernel(uint64_t seed, uint32_t per_thread, unsigned long long *mismatches)
{
    const uint64_t tid = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    unsigned long long bad = 0;
    for (uint32_t i = 0; i < per_thread; ++i)
    {
        const uint64_t u = st_mix(seed ^ (tid * 0x100000001B3ull + i));
        const uint64_t w = st_mix(u);
        const uint64_t sel = st_mix(w);
        // divisor: exponent within ±60 of 1.0, mantissa random / all ones / all zeros / one bit
        uint64_t sm = u & 0xFFFFFFFFFFFFFull;
        switch (sel & 7u)
        {
        case 0: sm = 0xFFFFFFFFFFFFFull; break;
        case 1: sm = 0; break;
        case 2: sm = 1ull << (sel >> 8) % 52; break;
        case 3: sm = 0xFFFFFFFFFFFFFull ^ (1ull << (sel >> 8) % 52); break;
        default: break;
        }
        const uint64_t se = 1023 - 60 + (u >> 52) % 121;
        const double s = __longlong_as_double(static_cast<long long>((se << 52) | sm));
        double a;
        if ((sel >> 3) & 1u)
        {
            // a = RN(q·s) moved by −2..+2 ulp: quotient next to the representable q
            const uint64_t qe = 1023 - 60 + (w >> 52) % 121;
            const double q = __longlong_as_double(static_cast<long long>((qe << 52) | (w & 0xFFFFFFFFFFFFFull)));
            const double p = __dmul_rn(q, s);
            const long long d = static_cast<long long>((sel >> 16) % 5) - 2;
            a = __longlong_as_double(__double_as_longlong(p) + d);
        }
        else
        {
            const uint64_t ae = 1023 - 60 + (w >> 52) % 121;
            a = __longlong_as_double(static_cast<long long>((ae << 52) | (w & 0xFFFFFFFFFFFFFull)));
        }
        if ((sel >> 4) & 1u) a = -a;
        const double want = a / s;
        const double got = pk_div_by_rcp(a, s, __drcp_rn(s));
        if (__double_as_longlong(want) != __double_as_longlong(got)) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// NVTX range over the host side of a call (the spans a timeline tool shows above the kernels they
// enqueue; header-only NVTX 3, a no-op without a profiler attached)
struct PkRange
{
    explicit PkRange(const char *name) { nvtxRangePushA(name); }
    ~PkRange() { nvtxRangePop(); }
    PkRange(const PkRange &) = delete;
    PkRange &operator=(const PkRange &) = delete;
};

#define PK_CUDA(call)                                                                                     \
    do                                                                                                    \
    {                                                                                                     \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
        {                                                                                                 \
            ctx->last_error = std::string(#call) + ": " + cudaGetErrorString(e__);                        \
            return (e__ == cudaErrorMemoryAllocation) ? PK_E_OOM : PK_E_CUDA;                             \
        }                                                                                                 \
    } while (0)

template <typename T> int dev_alloc(pk_ctx *ctx, T **p, size_t count)
{
    if (count == 0) count = 1;
    PK_CUDA(cudaMalloc(reinterpret_cast<void **>(p), count * sizeof(T)));
    return PK_OK;
}
#define PK_TRY(expr)              \
    do                            \
    {                             \
        int s__ = (expr);         \
        if (s__ != PK_OK) return s__; \
    } while (0)

inline uint32_t div_up(uint64_t a, uint64_t b) { return static_cast<uint32_t>((a + b - 1) / b); }

int bits_for(uint64_t n) // bits needed to represent values < n
{
    int b = 1;
    while (b < 64 && (1ull << b) < n) ++b;
    return b;
}

// LSD radix sort over the listed byte shifts; returns the index (0/1) of the buffer holding the result.
// n_dev != nullptr: the element count lives on the device (≤ n, for which the grid is sized).
int radix_sort(pk_ctx *ctx, uint64_t *keys[2], uint32_t *vals[2], uint64_t n, const std::vector<int> &shifts, int *result, int lowbits = 0,
               const unsigned long long *n_dev = nullptr)
{
    int cur = 0;
    uint32_t ntiles = div_up(n, SORT_TILE);
    if (ntiles <= 1 && shifts.size() <= 8 && !ctx->no_tile_sort)
    {
        // small worlds: all passes in one launch of one block (radix_sort_tile_kernel)
        uint64_t packed = 0;
        for (size_t k = 0; k < shifts.size(); ++k) packed |= static_cast<uint64_t>(shifts[k] & 0xFF) << (8 * k);
        if (vals)
            radix_sort_tile_kernel<true><<<1, SORT_THREADS, 0, ctx->stream>>>(keys[0], vals[0], keys[1], vals[1], n, n_dev, packed,
                                                                            static_cast<int>(shifts.size()), lowbits);
        else
            radix_sort_tile_kernel<false><<<1, SORT_THREADS, 0, ctx->stream>>>(keys[0], nullptr, keys[1], nullptr, n, n_dev, packed,
                                                                             static_cast<int>(shifts.size()), lowbits);
        ctx->launches += 1;
        PK_CUDA(cudaGetLastError());
        *result = static_cast<int>(shifts.size() & 1);
        return PK_OK;
    }
    if (static_cast<size_t>(ntiles) * 256 > ctx->tile_hist_entries)
    {
        ctx->last_error = "radix sort scratch too small";
        return PK_E_STATE;
    }
    for (int shift : shifts)
    {
        radix_hist_kernel<<<ntiles, SORT_THREADS, 0, ctx->stream>>>(keys[cur], n, n_dev, shift, lowbits, ctx->d_tile_hist, ntiles);
        radix_scan_kernel<<<256, SORT_THREADS, 0, ctx->stream>>>(ctx->d_tile_hist, ntiles, ctx->d_digit_total);
        if (vals)
            radix_scatter_kernel<true><<<ntiles, SORT_THREADS, 0, ctx->stream>>>(
                keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, n_dev, shift, lowbits, ctx->d_tile_hist, ntiles, ctx->d_digit_total);
        else
            radix_scatter_kernel<false><<<ntiles, SORT_THREADS, 0, ctx->stream>>>(
                keys[cur], nullptr, keys[cur ^ 1], nullptr, n, n_dev, shift, lowbits, ctx->d_tile_hist, ntiles, ctx->d_digit_total);
        ctx->launches += 3;
        cur ^= 1;
    }
    PK_CUDA(cudaGetLastError());
    *result = cur;
    return PK_OK;
}

int upload_shapes(pk_ctx *ctx)
{
    if (!ctx->shapes_dirty) return PK_OK;
    if (!ctx->h_shapes.empty())
        PK_CUDA(cudaMemcpyAsync(ctx->d_shapes, ctx->h_shapes.data(), ctx->h_shapes.size() * sizeof(ShapeRec),
                                cudaMemcpyHostToDevice, ctx->stream));
    std::vector<float4> vf;
    if (!ctx->h_verts.empty())
    {
        PK_CUDA(cudaMemcpyAsync(ctx->d_verts, ctx->h_verts.data(), ctx->h_verts.size() * sizeof(double),
                                cudaMemcpyHostToDevice, ctx->stream));
        // float copy for the hull-support prefilter (round to nearest; the error bound covers it)
        vf.resize(ctx->h_verts.size() / 3);
        for (size_t i = 0; i < vf.size(); ++i)
            vf[i] = make_float4(static_cast<float>(ctx->h_verts[3 * i]), static_cast<float>(ctx->h_verts[3 * i + 1]),
                                static_cast<float>(ctx->h_verts[3 * i + 2]), 0.f);
        PK_CUDA(cudaMemcpyAsync(ctx->d_verts_f, vf.data(), vf.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    }
    PK_CUDA(cudaStreamSynchronize(ctx->stream)); // host vectors may be reallocated by the next pk_shape_*
    ctx->shapes_dirty = false;
    return PK_OK;
}

int add_shape(pk_ctx *ctx, const ShapeRec &r, uint32_t *id)
{
    if (!ctx || !id) return PK_E_INVALID;
    if (ctx->h_shapes.size() >= ctx->cfg.max_shapes)
    {
        ctx->last_error = "shape table full (pk_config.max_shapes)";
        return PK_E_INVALID;
    }
    *id = static_cast<uint32_t>(ctx->h_shapes.size());
    ctx->h_shapes.push_back(r);
    ctx->shapes_dirty = true;
    return PK_OK;
}

BodyArrays body_arrays(pk_ctx *ctx)
{
    BodyArrays b;
    b.shapes = ctx->d_shapes;
    b.verts = ctx->d_verts;
    b.verts_f = ctx->d_verts_f;
    b.pos = ctx->d_pos;
    b.quat = ctx->d_quat;
    b.shape_id = ctx->d_shape_id;
    return b;
}

// GJK → scan → EPA over npairs pairs given either as sorted keys or as explicit index arrays.
// Leaves contacts (slot order = pair order among GJK hits) in ctx->d_contacts[0], validity in d_valid.
// npairs_dev != nullptr: the pair count lives on the device (≤ npairs, for which the grids are sized).
int run_narrowphase(pk_ctx *ctx, const uint64_t *d_keys, const uint32_t *d_a, const uint32_t *d_b, uint64_t npairs,
                    bool timed, ContactRec *mirror = nullptr, const unsigned long long *npairs_dev = nullptr)
{
    PkRange range("pk: narrowphase (GJK, hit scan, EPA)");
    if (timed) cudaEventRecord(ctx->ev[ST_GJK], ctx->stream);
    if (npairs)
    {
        const uint32_t pair_blocks = div_up(npairs, 128);
        auto gjk = [&](auto kernel, const GjkCarry *carry)
        {
            kernel<<<div_up(npairs, PK_GJK_THREADS), PK_GJK_THREADS, 0, ctx->stream>>>(
                body_arrays(ctx), d_keys, d_a, d_b, ctx->d_gjk_work, ctx->cfg.max_pairs, ctx->d_counters + C_GJK_CLASS, ctx->d_hit, ctx->d_simplices,
                ctx->d_counters + C_HITS, ctx->max_contacts, ctx->d_counters + C_CLASS_COUNT, carry);
        };
        if (!ctx->exact_prefilter && !ctx->has_big_hulls)
        {
            // misses are settled in FP32 with a certificate (pk_gjk_filter.cuh); what is left runs the reference's
            // iteration from its first support
            gjk_filter_kernel<<<pair_blocks, 128, 0, ctx->stream>>>(body_arrays(ctx), d_keys, d_a, d_b, npairs, npairs_dev, ctx->d_hit, ctx->d_gjk_work,
                                                                    ctx->cfg.max_pairs, ctx->d_counters + C_GJK_CLASS, ctx->filter_iters);
            gjk(gjk_kernel<false, false>, nullptr);
        }
        else if (ctx->has_big_hulls)
        {
            // many-vertex hulls: the reference's first two supports for every pair in FP64, carried to gjk_kernel (see
            // gjk_prefilter_kernel); pairs of analytic shapes are filtered in FP32 all the same
            if (!ctx->d_gjk_carry) PK_TRY(dev_alloc(ctx, &ctx->d_gjk_carry, ctx->cfg.max_pairs)); // 96 B per pair of capacity
            auto pre = ctx->exact_prefilter ? gjk_prefilter_kernel<true, false> : gjk_prefilter_kernel<true, true>;
            pre<<<pair_blocks, 128, 0, ctx->stream>>>(body_arrays(ctx), d_keys, d_a, d_b, npairs, npairs_dev, ctx->d_hit, ctx->d_gjk_work, ctx->cfg.max_pairs,
                                                      ctx->d_counters + C_GJK_CLASS, ctx->d_gjk_carry);
            gjk(gjk_kernel<true, true>, ctx->d_gjk_carry);
        }
        else
        {
            // PK_GJK_EXACT_PREFILTER=1: round 1's form, kept for A/B runs
            gjk_prefilter_kernel<false><<<pair_blocks, 128, 0, ctx->stream>>>(body_arrays(ctx), d_keys, d_a, d_b, npairs, npairs_dev, ctx->d_hit, ctx->d_gjk_work,
                                                                              ctx->cfg.max_pairs, ctx->d_counters + C_GJK_CLASS, nullptr);
            gjk(gjk_kernel<false, false>, nullptr);
        }
        ctx->launches += 2;
    }
    if (timed) cudaEventRecord(ctx->ev[ST_SCAN], ctx->stream);
    if (npairs)
    {
        uint32_t nt = div_up(npairs, SCAN_TILE);
        flag_tile_sum_kernel<<<nt, 256, 0, ctx->stream>>>(ctx->d_hit, npairs, npairs_dev, ctx->d_scan_tiles);
        tile_sum_scan_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d_scan_tiles, nt, ctx->d_counters + C_SCAN_TOTAL);
        flag_scan_apply_kernel<<<nt, 256, 0, ctx->stream>>>(ctx->d_hit, npairs, npairs_dev, ctx->d_scan_tiles, ctx->d_out_index);
        ctx->launches += 3;
    }
    if (timed) cudaEventRecord(ctx->ev[ST_EPA], ctx->stream);
    if (npairs)
    {
        epa_order_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->d_simplices, ctx->d_counters + C_HITS, ctx->max_contacts,
                                                                     ctx->d_counters + C_CLASS_COUNT, ctx->d_counters + C_CLASS_FILL,
                                                                     ctx->d_epa_order);
        // One persistent launch for all hits (pk_epa_coop.cuh); what it hands back (padded simplices, improper
        // horizons, polytopes past the slab: a handful) is left to epa_kernel.
        {
            const uint64_t most = std::min<uint64_t>(npairs, ctx->max_contacts);
            epa_init_kernel<<<div_up(most, 128), 128, 0, ctx->stream>>>(ctx->d_simplices, ctx->d_counters + C_HITS, ctx->max_contacts, d_keys,
                                                                       d_a, d_b, ctx->d_out_index, ctx->d_epa_init);
        }
        EcParams ep;
        ep.bodies = body_arrays(ctx);
        ep.simplices = ctx->d_simplices;
        ep.hit_count_ptr = ctx->d_counters + C_HITS;
        ep.hit_capacity = ctx->max_contacts;
        ep.order = ctx->d_epa_order;
        ep.contacts = ctx->d_contacts[0];
        ep.valid = ctx->d_valid;
        ep.slabs = ctx->d_epa_spill;
        ep.cursor = ctx->d_counters + C_EPA_SCAN_CURSOR;
        ep.counters = ctx->d_counters + C_VALID;
        ep.fallback = ctx->d_epa_fallback2;
        ep.fallback_count = ctx->d_counters + C_EPA_FALLBACK + 1;
        ep.restart_count = ctx->d_counters + C_EPA_FALLBACK;
        ep.init = ctx->d_epa_init;
        ep.contacts_host = mirror;
        auto coop = ctx->has_big_hulls ? (mirror ? epa_coop_kernel<true, true> : epa_coop_kernel<false, true>)
                                       : (mirror ? epa_coop_kernel<true, false> : epa_coop_kernel<false, false>);
        coop<<<ctx->epa_scan_blocks, ES_THREADS, 0, ctx->stream>>>(ep);
        epa_kernel<<<ctx->epa_blocks, EPA_THREADS, 0, ctx->stream>>>(
            body_arrays(ctx), d_keys, d_a, d_b, ctx->d_simplices, ctx->d_counters + C_EPA_FALLBACK + 1, ctx->max_contacts,
            ctx->d_out_index, ctx->d_epa_fallback2, ctx->d_contacts[0], ctx->d_valid, ctx->d_slabs,
            ctx->d_counters + C_EPA_FB_CURSOR, ctx->d_counters + C_VALID, mirror);
        ctx->launches += 4;
    }
    if (timed) cudaEventRecord(ctx->ev[ST_COMPACT], ctx->stream);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

int ensure_host_pairs(pk_ctx *ctx, uint64_t n)
{
    if (n <= ctx->h_pairs_cap) return PK_OK;
    if (ctx->h_pairs) cudaFreeHost(ctx->h_pairs);
    ctx->h_pairs = nullptr;
    ctx->h_pairs_cap = 0;
    size_t cap = std::min<size_t>(ctx->cfg.max_pairs, std::max<size_t>(n * 5 / 4, 1024));
    PK_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_pairs), cap * sizeof(uint64_t), cudaHostAllocDefault));
    ctx->h_pairs_cap = cap;
    return PK_OK;
}

int read_counters(pk_ctx *ctx)
{
    PK_CUDA(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, C_COUNT * sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    return PK_OK;
}

// The step can be repeated after pk_reserve_pairs: the pair set is a function of the per-body state (stored box,
// last_move, create), which this attempt has already brought up to date for this epoch; a second pass over the
// same poses in the same epoch leaves it as it is (every true box lies inside its stored box now), so only
// num_moved of the repeated step differs (0).
int pair_overflow(pk_ctx *ctx, pk_step_result *out, uint64_t npairs)
{
    if (out)
    {
        std::memset(out, 0, sizeof(*out));
        out->pairs_required = npairs;
        out->num_moved = ctx->h_counters[C_MOVED];
        out->step_index = static_cast<uint32_t>(ctx->epoch);
    }
    ctx->last_error = "candidate pairs exceed pk_config.max_pairs (pk_reserve_pairs, then repeat the step)";
    return PK_E_PAIR_OVERFLOW;
}

// The same step once more, with the pair count read back in the middle and (rows_off) the pairs as a sorted list: the
// pair set is a function of the per-body state, which the first pass has brought up to date (see pair_overflow), so the
// second pass finds the same pairs; only its num_moved (0) is not the step's.
int repeat_step(pk_ctx *ctx, pk_step_result *out, bool rows_off)
{
    const unsigned long long moved = ctx->h_counters[C_MOVED];
    if (rows_off) ctx->rows_off = true;
    ctx->pairs_read_back = true;
    const int rc = pk_collide_resident(ctx, out);
    ctx->pairs_read_back = false;
    if (out) out->num_moved = moved;
    return rc;
}

} // namespace

extern "C"
{

int pk_abi_version(void) { return PK_ABI_VERSION; }

const char *pk_strerror(int status)
{
    switch (status)
    {
    case PK_OK: return "ok";
    case PK_E_INVALID: return "invalid argument";
    case PK_E_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
    case PK_E_CUDA: return "CUDA runtime error";
    case PK_E_OOM: return "out of device or pinned memory";
    case PK_E_PAIR_OVERFLOW: return "candidate pair / contact capacity exceeded";
    case PK_E_EPA_OVERFLOW: return "EPA polytope scratch exceeded";
    case PK_E_STATE: return "call order violated";
    default: return "unknown status";
    }
}

const char *pk_last_error(pk_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }

// Everything sized by pk_config.max_pairs / max_contacts (pk_create, pk_reserve_pairs).
static int alloc_pair_buffers(pk_ctx *ctx)
{
    const size_t nb = ctx->cfg.max_bodies;
    const size_t np = ctx->cfg.max_pairs;
    const size_t nc = ctx->max_contacts;
    int s;
#define A(ptr, count)                                   \
    if ((s = dev_alloc(ctx, &(ptr), (count))) != PK_OK) \
    return s
    ctx->tile_hist_entries = static_cast<size_t>(div_up(std::max(nb, np), SORT_TILE)) * 256;
    A(ctx->d_tile_hist, ctx->tile_hist_entries);
    A(ctx->d_digit_total, 256);
    A(ctx->d_pkeys[0], np);
    A(ctx->d_pkeys[1], np);
    A(ctx->d_hit, np + 16);
    A(ctx->d_out_index, np);
    A(ctx->d_scan_tiles, div_up(np, SCAN_TILE) + 1);
    A(ctx->d_simplices, nc);
    A(ctx->d_contacts[0], nc);
    A(ctx->d_contacts[1], nc);
    A(ctx->d_valid, nc + 16);
    A(ctx->d_valid_index, nc);
    A(ctx->d_epa_order, nc);
    A(ctx->d_gjk_work, 4 * np); // one survivor list per shape-kind class
    // persistent EPA grid: enough resident threads to fill the machine, never more than the work
    {
        int gjk_per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&gjk_per_sm, gjk_kernel<false>, PK_GJK_THREADS, 0) != cudaSuccess || gjk_per_sm < 1)
            gjk_per_sm = 2;
        ctx->gjk_blocks = static_cast<uint32_t>(ctx->sm_count * gjk_per_sm);
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, epa_kernel, EPA_THREADS, 0) != cudaSuccess || per_sm < 1)
            per_sm = 4;
#ifdef PK_EPA_BLOCKS_PER_SM
        per_sm = std::min(per_sm, PK_EPA_BLOCKS_PER_SM);
#endif
        uint64_t want_threads = static_cast<uint64_t>(ctx->sm_count) * per_sm * EPA_THREADS;
        uint64_t need_threads = ((nc + EPA_THREADS - 1) / EPA_THREADS) * EPA_THREADS;
        uint64_t threads = std::min(want_threads, std::max<uint64_t>(need_threads, EPA_THREADS));
        ctx->epa_blocks = static_cast<uint32_t>(threads / EPA_THREADS);
        A(ctx->d_slabs, threads * EPA_SLAB_BYTES);
        // epa_coop_kernel: 43 KB of shared memory per 64-thread block, as many blocks per SM as fit
        cudaFuncSetAttribute(epa_coop_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(epa_coop_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(epa_coop_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(epa_coop_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int es_per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&es_per_sm, epa_coop_kernel<false, true>, ES_THREADS, 0) != cudaSuccess || es_per_sm < 1)
            es_per_sm = 1;
#ifdef PK_ES_BLOCKS_PER_SM
        es_per_sm = std::min(es_per_sm, PK_ES_BLOCKS_PER_SM);
#endif
#ifdef PK_ES_FORCE_BLOCKS
        es_per_sm = PK_ES_FORCE_BLOCKS;
#endif
        if (getenv("PK_DEBUG")) fprintf(stderr, "[pk] epa_coop_kernel: %d blocks per SM\n", es_per_sm);
        const uint64_t es_need = (nc + ES_THREADS - 1) / ES_THREADS;
        ctx->epa_scan_blocks = static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(static_cast<uint64_t>(ctx->sm_count) * es_per_sm, es_need)));
        A(ctx->d_epa_spill, static_cast<size_t>(ctx->epa_scan_blocks) * ES_THREADS * es_slab_bytes());
        A(ctx->d_epa_fallback2, nc);
        A(ctx->d_epa_init, nc);
    }
#undef A
    cudaMemsetAsync(ctx->d_hit, 0, np + 16, ctx->stream);
    return PK_OK;
}

static void free_pair_buffers(pk_ctx *ctx)
{
    for (void **q : {reinterpret_cast<void **>(&ctx->d_tile_hist), reinterpret_cast<void **>(&ctx->d_digit_total),
                     reinterpret_cast<void **>(&ctx->d_pkeys[0]), reinterpret_cast<void **>(&ctx->d_pkeys[1]),
                     reinterpret_cast<void **>(&ctx->d_hit), reinterpret_cast<void **>(&ctx->d_out_index),
                     reinterpret_cast<void **>(&ctx->d_scan_tiles), reinterpret_cast<void **>(&ctx->d_simplices),
                     reinterpret_cast<void **>(&ctx->d_contacts[0]), reinterpret_cast<void **>(&ctx->d_contacts[1]),
                     reinterpret_cast<void **>(&ctx->d_valid), reinterpret_cast<void **>(&ctx->d_valid_index),
                     reinterpret_cast<void **>(&ctx->d_epa_order), reinterpret_cast<void **>(&ctx->d_gjk_work),
                     reinterpret_cast<void **>(&ctx->d_gjk_carry),
                     reinterpret_cast<void **>(&ctx->d_slabs), reinterpret_cast<void **>(&ctx->d_epa_spill),
                     reinterpret_cast<void **>(&ctx->d_epa_fallback2), reinterpret_cast<void **>(&ctx->d_epa_init)})
    {
        if (*q) cudaFree(*q);
        *q = nullptr;
    }
    ctx->d_pairs_sorted = nullptr;
    ctx->d_contacts_final = nullptr;
}

int pk_destroy(pk_ctx *ctx)
{
    if (!ctx) return PK_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
#ifdef PK_EC_TIMING
    {
        static unsigned long long h[8192][2];
        if (cudaMemcpyFromSymbol(h, g_ec_times, sizeof(h)) == cudaSuccess)
            for (int w = 0; w < ctx->epa_scan_blocks * (ES_THREADS / 32); ++w) printf("[ec] %d %llu %llu\n", w, h[w][0], h[w][1]);
    }
#endif
    void *dev[] = {ctx->d_verts_f, ctx->d_shapes,      ctx->d_verts,        ctx->d_pos,         ctx->d_quat,         ctx->d_disp,
                   ctx->d_shape_id,    ctx->d_world,        ctx->d_flags,       ctx->st.stored,      ctx->st.last_move,
                   ctx->st.create,     ctx->st.alive,       ctx->d_scene,       ctx->d_counters,     ctx->d_bkeys[0],
                   ctx->d_bkeys[1],    ctx->d_bvals[0],     ctx->d_bvals[1],    ctx->d_tile_hist,    ctx->d_digit_total,
                   ctx->d_leaves,      ctx->d_nodes,        ctx->d_right,       ctx->d_range_last,   ctx->d_root,
                   ctx->d_merge_flag,  ctx->d_pkeys[0],     ctx->d_pkeys[1],    ctx->d_hit,          ctx->d_out_index,
                   ctx->d_scan_tiles,  ctx->d_simplices,    ctx->d_contacts[0], ctx->d_contacts[1],  ctx->d_valid,
                   ctx->d_valid_index, ctx->d_slabs,       ctx->d_epa_order,    ctx->d_gjk_work,
                   ctx->d_epa_spill,   ctx->d_epa_fallback2, ctx->d_epa_init,     ctx->d_gjk_carry,
                   ctx->d_row_count,   ctx->d_rows,         ctx->d_row_tiles};
    for (void *p : dev)
        if (p) cudaFree(p);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->h_pairs) cudaFreeHost(ctx->h_pairs);
    if (ctx->h_contacts) cudaFreeHost(ctx->h_contacts);
    {
        void *man[] = {ctx->d_man[0], ctx->d_man[1], ctx->d_man_stage, ctx->d_man_ckeys[0], ctx->d_man_ckeys[1], ctx->d_man_csrc[0],
                       ctx->d_man_csrc[1], ctx->d_man_began[0], ctx->d_man_began[1], ctx->d_man_ended[0], ctx->d_man_ended[1],
                       ctx->d_man_consumed, ctx->d_man_counters, ctx->d_man_imp};
        for (void *q : man)
            if (q) cudaFree(q);
        if (ctx->h_man) cudaFreeHost(ctx->h_man);
        if (ctx->h_man_events) cudaFreeHost(ctx->h_man_events);
    }
    if (ctx->d_points) cudaFree(ctx->d_points);
    if (ctx->h_points) cudaFreeHost(ctx->h_points);
    for (void *q : {static_cast<void *>(ctx->d_ray_o), static_cast<void *>(ctx->d_ray_d), static_cast<void *>(ctx->d_ray_max),
                    static_cast<void *>(ctx->d_ray_world), static_cast<void *>(ctx->d_ray_keys[0]), static_cast<void *>(ctx->d_ray_keys[1]),
                    static_cast<void *>(ctx->d_ray_src[0]), static_cast<void *>(ctx->d_ray_src[1]), static_cast<void *>(ctx->d_ray_dist),
                    static_cast<void *>(ctx->d_ray_hits), static_cast<void *>(ctx->d_ray_counter)})
        if (q) cudaFree(q);
    for (auto &e : ctx->ev_ray)
        if (e) cudaEventDestroy(e);
    for (void *q : {static_cast<void *>(ctx->dyn.vel), static_cast<void *>(ctx->dyn.ang_vel), static_cast<void *>(ctx->dyn.acc),
                    static_cast<void *>(ctx->dyn.torque), static_cast<void *>(ctx->dyn.mass), static_cast<void *>(ctx->dyn.inertia),
                    static_cast<void *>(ctx->dyn.inertia_w)})
        if (q) cudaFree(q);
    for (auto &e : ctx->ev_dyn)
        if (e) cudaEventDestroy(e);
    for (void *q : {static_cast<void *>(ctx->d_material), static_cast<void *>(ctx->d_sol_valid), static_cast<void *>(ctx->d_sol_index),
                    static_cast<void *>(ctx->d_sol_tiles), static_cast<void *>(ctx->d_sol_slot), static_cast<void *>(ctx->d_sol_rows), static_cast<void *>(ctx->d_sol_total)})
        if (q) cudaFree(q);
    if (ctx->h_sol_rows) cudaFreeHost(ctx->h_sol_rows);
    for (auto &e : ctx->ev_sol)
        if (e) cudaEventDestroy(e);
    for (auto &e : ctx->ev)
        if (e) cudaEventDestroy(e);
    if (ctx->copy_stream)
    {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    if (ctx->comm) nccl_api().CommDestroy(ctx->comm);
    if (ctx->d_comm_counts) cudaFree(ctx->d_comm_counts);
    if (ctx->h_comm_counts) cudaFreeHost(ctx->h_comm_counts);
    if (ctx->d_gather) cudaFree(ctx->d_gather);
    for (auto &e : ctx->ev_comm)
        if (e) cudaEventDestroy(e);
    if (ctx->ev_sorted) cudaEventDestroy(ctx->ev_sorted);
    if (ctx->ev_pairs_copied) cudaEventDestroy(ctx->ev_pairs_copied);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return PK_OK;
}

int pk_create(const pk_config *cfg, pk_ctx **out)
{
    if (!cfg || !out) return PK_E_INVALID;
    *out = nullptr;
    if (cfg->max_bodies == 0 || cfg->max_pairs == 0 || cfg->max_shapes == 0) return PK_E_INVALID;
    if (cfg->mode != PK_MODE_WORLD && cfg->mode != PK_MODE_QUERY) return PK_E_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev)
        return PK_E_NO_DEVICE;
    if (cudaSetDevice(cfg->device) != cudaSuccess) return PK_E_NO_DEVICE;
    pk_ctx *ctx = new (std::nothrow) pk_ctx();
    if (!ctx) return PK_E_OOM;
    ctx->cfg = *cfg;
    if (ctx->cfg.num_worlds == 0) ctx->cfg.num_worlds = 1;
    if (ctx->cfg.shard_count == 0) ctx->cfg.shard_count = 1;
    if (ctx->cfg.shard_rank >= ctx->cfg.shard_count)
    {
        delete ctx;
        return PK_E_INVALID;
    }
    ctx->max_contacts = cfg->max_contacts ? cfg->max_contacts : cfg->max_pairs;
    ctx->exact_prefilter = getenv("PK_GJK_EXACT_PREFILTER") != nullptr;
    if (const char *e = getenv("PK_GJK_FILTER_ITERS")) ctx->filter_iters = atoi(e);
    ctx->no_tile_sort = getenv("PK_NO_TILE_SORT") != nullptr;
    auto fail = [&](int s)
    {
        pk_destroy(ctx);
        return s;
    };
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(PK_E_NO_DEVICE);
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(PK_E_CUDA);
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return fail(PK_E_CUDA);
    if (cudaEventCreateWithFlags(&ctx->ev_sorted, cudaEventDisableTiming) != cudaSuccess) return fail(PK_E_CUDA);
    if (cudaEventCreateWithFlags(&ctx->ev_pairs_copied, cudaEventDisableTiming) != cudaSuccess) return fail(PK_E_CUDA);
    for (auto &e : ctx->ev)
        if (cudaEventCreate(&e) != cudaSuccess) return fail(PK_E_CUDA);

    const size_t nb = cfg->max_bodies;
    const size_t np = cfg->max_pairs;
    int s;
#define A(ptr, count)                                   \
    if ((s = dev_alloc(ctx, &(ptr), (count))) != PK_OK) \
    return fail(s)
    A(ctx->d_shapes, cfg->max_shapes);
    A(ctx->d_verts, (cfg->max_hull_vertices + cfg->max_shapes) * 3 + 2); // + one padding vertex per hull at most
    A(ctx->d_verts_f, cfg->max_hull_vertices + cfg->max_shapes + 2);
    A(ctx->d_pos, nb * 3);
    A(ctx->d_quat, nb * 4);
    A(ctx->d_disp, nb * 3);
    A(ctx->d_shape_id, nb);
    A(ctx->d_world, nb);
    A(ctx->d_flags, nb);
    A(ctx->st.stored, nb * 6);
    A(ctx->st.last_move, nb);
    A(ctx->st.create, nb);
    A(ctx->st.alive, nb);
    A(ctx->d_scene, 8);
    A(ctx->d_counters, C_COUNT);
    A(ctx->d_bkeys[0], nb);
    A(ctx->d_bkeys[1], nb);
    A(ctx->d_bvals[0], nb);
    A(ctx->d_bvals[1], nb);
    A(ctx->d_leaves, nb);
    A(ctx->d_nodes, 2 * nb);
    A(ctx->d_right, nb);
    A(ctx->d_range_last, nb);
    A(ctx->d_root, 1);
    A(ctx->d_merge_flag, nb);
    A(ctx->d_row_count, nb);
    A(ctx->d_rows, nb * PAIR_ROW);
    A(ctx->d_row_tiles, div_up(nb, ROWS_TILE) + 1);
    ctx->rows_off = getenv("PK_PAIR_RADIX") != nullptr;
    if ((s = alloc_pair_buffers(ctx)) != PK_OK) return fail(s);
#undef A
    if (cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_counters), C_COUNT * sizeof(unsigned long long),
                      cudaHostAllocDefault) != cudaSuccess)
        return fail(PK_E_OOM);
    cudaMemsetAsync(ctx->st.alive, 0, nb, ctx->stream);
    cudaMemsetAsync(ctx->d_flags, 0, nb, ctx->stream);
    cudaMemsetAsync(ctx->d_world, 0, nb * sizeof(uint32_t), ctx->stream);
    cudaMemsetAsync(ctx->d_disp, 0, nb * 3 * sizeof(double), ctx->stream);
    cudaMemsetAsync(ctx->d_hit, 0, np + 16, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(PK_E_CUDA);
    ctx->h_flags.assign(nb, 0);
    *out = ctx;
    return PK_OK;
}

int pk_reserve_pairs(pk_ctx *ctx, uint64_t max_pairs, uint64_t max_contacts)
{
    if (!ctx || max_pairs == 0) return PK_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    if (max_contacts == 0) max_contacts = max_pairs;
    if (max_pairs <= ctx->cfg.max_pairs && max_contacts <= ctx->max_contacts) return PK_OK; // never shrinks
    max_pairs = std::max<uint64_t>(max_pairs, ctx->cfg.max_pairs);
    max_contacts = std::max<uint64_t>(max_contacts, ctx->max_contacts);
    if (ctx->pairs_in_flight) cudaStreamSynchronize(ctx->copy_stream);
    ctx->pairs_in_flight = false;
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    free_pair_buffers(ctx);
    ctx->cfg.max_pairs = max_pairs;
    ctx->max_contacts = max_contacts;
    ctx->cfg.max_contacts = max_contacts;
    ctx->have_results = false; // the last step's device results lived in the old buffers
    ctx->device_results = false;
    ctx->fetched = false;
    PK_TRY(alloc_pair_buffers(ctx));
    if (ctx->d_man_consumed)
    {
        cudaFree(ctx->d_man_consumed);
        ctx->d_man_consumed = nullptr;
        PK_TRY(dev_alloc(ctx, &ctx->d_man_consumed, ctx->max_contacts + 16));
    }
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    return PK_OK;
}

// ------------------------------------------------------------------------------------ shapes
int pk_shape_box(pk_ctx *ctx, const double half[3], uint32_t *id)
{
    if (!ctx || !half) return PK_E_INVALID;
    ShapeRec r{};
    r.kind = KIND_OBB;
    for (int k = 0; k < 3; ++k) r.a[k] = half[k];
    return add_shape(ctx, r, id);
}

int pk_shape_sphere(pk_ctx *ctx, double radius, uint32_t *id)
{
    if (!ctx) return PK_E_INVALID;
    ShapeRec r{};
    r.kind = KIND_SPHERE;
    r.a[0] = radius;
    return add_shape(ctx, r, id);
}

int pk_shape_aabb(pk_ctx *ctx, const double mn[3], const double mx[3], uint32_t *id)
{
    if (!ctx || !mn || !mx) return PK_E_INVALID;
    ShapeRec r{};
    r.kind = KIND_AABB;
    for (int k = 0; k < 3; ++k) r.a[k] = mn[k], r.b[k] = mx[k];
    return add_shape(ctx, r, id);
}

int pk_shape_hull(pk_ctx *ctx, const double *xyz, uint32_t nverts, uint32_t *id)
{
    if (!ctx || !xyz || nverts == 0) return PK_E_INVALID;
    size_t off = ctx->h_verts.size() / 3;
    if (off & 1u) // hulls start at even vertex offsets: the float scan loads vertex pairs with 32-byte loads
    {
        ctx->h_verts.insert(ctx->h_verts.end(), 3, 0.0);
        ++off;
    }
    if (off + nverts > ctx->cfg.max_hull_vertices + ctx->cfg.max_shapes)
    {
        ctx->last_error = "hull vertex pool full (pk_config.max_hull_vertices)";
        return PK_E_INVALID;
    }
    ShapeRec r{};
    r.kind = KIND_HULL;
    r.vert_off = static_cast<uint32_t>(off);
    r.nverts = nverts;
    // aabb::from_points (bounds.h:33-49): the mesh-local box that instance::bounds() rotates
    for (int k = 0; k < 3; ++k) r.a[k] = r.b[k] = xyz[k];
    for (uint32_t i = 1; i < nverts; ++i)
        for (int k = 0; k < 3; ++k)
        {
            r.a[k] = std::min(r.a[k], xyz[3 * i + k]);
            r.b[k] = std::max(r.b[k], xyz[3 * i + k]);
        }
    // mesh::box's vertex table (src/mesh.cpp:31-40), bit for bit: support() then shares the eight dots' products
    if (nverts == 8 && r.b[0] > 0.0 && r.b[1] > 0.0 && r.b[2] > 0.0)
    {
        bool box = true;
        for (uint32_t i = 0; i < 8 && box; ++i)
        {
            const double want[3] = {((0x66 >> i) & 1) ? r.b[0] : -r.b[0], ((0xCC >> i) & 1) ? r.b[1] : -r.b[1], i >= 4 ? r.b[2] : -r.b[2]};
            for (int k = 0; k < 3; ++k) box = box && xyz[3 * i + k] == want[k];
        }
        r.mesh_box = box ? 1u : 0u;
    }
    int s = add_shape(ctx, r, id);
    if (s != PK_OK) return s;
    if (nverts > HULL_PREFILTER_MIN) ctx->has_big_hulls = true;
    ctx->h_verts.insert(ctx->h_verts.end(), xyz, xyz + 3ull * nverts);
    return PK_OK;
}

int pk_shapes_bulk(pk_ctx *ctx, const int32_t *kind, const double *par3, uint32_t n, uint32_t *first_id)
{
    if (!ctx || !kind || !par3) return PK_E_INVALID;
    if (ctx->h_shapes.size() + n > ctx->cfg.max_shapes)
    {
        ctx->last_error = "shape table full (pk_config.max_shapes)";
        return PK_E_INVALID;
    }
    if (first_id) *first_id = static_cast<uint32_t>(ctx->h_shapes.size());
    for (uint32_t i = 0; i < n; ++i)
    {
        ShapeRec r{};
        if (kind[i] == KIND_OBB)
        {
            r.kind = KIND_OBB;
            for (int k = 0; k < 3; ++k) r.a[k] = par3[3 * i + k];
        }
        else if (kind[i] == KIND_SPHERE)
        {
            r.kind = KIND_SPHERE;
            r.a[0] = par3[3 * i];
        }
        else
            return PK_E_INVALID;
        ctx->h_shapes.push_back(r);
    }
    ctx->shapes_dirty = true;
    return PK_OK;
}

// ------------------------------------------------------------------------------------ bodies
int pk_bodies_resize(pk_ctx *ctx, uint32_t n)
{
    if (!ctx || n > ctx->cfg.max_bodies) return PK_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    if (n < ctx->n_bodies)
    {
        // bodies beyond n are destroyed
        std::fill(ctx->h_flags.begin() + n, ctx->h_flags.begin() + ctx->n_bodies, 0);
        PK_CUDA(cudaMemsetAsync(ctx->d_flags + n, 0, ctx->n_bodies - n, ctx->stream));
        PK_CUDA(cudaMemsetAsync(ctx->st.alive + n, 0, ctx->n_bodies - n, ctx->stream));
    }
    ctx->n_bodies = n;
    ctx->alive_dirty = true;
    return PK_OK;
}

int pk_bodies_upload(pk_ctx *ctx, const double *pos, const double *quat, const double *disp, const uint32_t *shape_id,
                     const uint8_t *flags, const uint32_t *world_id, uint32_t first, uint32_t count)
{
    if (!ctx || !pos || !quat || !shape_id || !flags) return PK_E_INVALID;
    if (static_cast<uint64_t>(first) + count > ctx->n_bodies) return PK_E_INVALID;
    if (count == 0) return PK_OK;
    cudaSetDevice(ctx->cfg.device);
    for (uint32_t i = 0; i < count; ++i)
        if ((flags[i] & FLAG_ALIVE) && shape_id[i] >= ctx->h_shapes.size())
        {
            ctx->last_error = "shape id out of range";
            return PK_E_INVALID;
        }
    cudaStream_t s = ctx->stream;
    PK_CUDA(cudaMemcpyAsync(ctx->d_pos + 3ull * first, pos, 3ull * count * sizeof(double), cudaMemcpyHostToDevice, s));
    PK_CUDA(cudaMemcpyAsync(ctx->d_quat + 4ull * first, quat, 4ull * count * sizeof(double), cudaMemcpyHostToDevice, s));
    if (disp)
        PK_CUDA(cudaMemcpyAsync(ctx->d_disp + 3ull * first, disp, 3ull * count * sizeof(double), cudaMemcpyHostToDevice, s));
    PK_CUDA(cudaMemcpyAsync(ctx->d_shape_id + first, shape_id, count * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    PK_CUDA(cudaMemcpyAsync(ctx->d_flags + first, flags, count, cudaMemcpyHostToDevice, s));
    if (world_id)
    {
        PK_CUDA(cudaMemcpyAsync(ctx->d_world + first, world_id, count * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        ctx->have_world = true;
    }
    std::memcpy(ctx->h_flags.data() + first, flags, count);
    ctx->alive_dirty = true;
    if (ctx->dyn_enabled) dynamics_derive_kernel<<<div_up(count, 256), 256, 0, s>>>(ctx->d_quat, ctx->dyn, first, count);
    return PK_OK;
}

int pk_bodies_update_pose(pk_ctx *ctx, const double *pos, const double *quat, const double *disp, uint32_t first,
                          uint32_t count)
{
    if (!ctx) return PK_E_INVALID;
    if (static_cast<uint64_t>(first) + count > ctx->n_bodies) return PK_E_INVALID;
    if (count == 0) return PK_OK;
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t s = ctx->stream;
    if (pos)
        PK_CUDA(cudaMemcpyAsync(ctx->d_pos + 3ull * first, pos, 3ull * count * sizeof(double), cudaMemcpyHostToDevice, s));
    if (quat)
        PK_CUDA(cudaMemcpyAsync(ctx->d_quat + 4ull * first, quat, 4ull * count * sizeof(double), cudaMemcpyHostToDevice, s));
    if (disp)
        PK_CUDA(cudaMemcpyAsync(ctx->d_disp + 3ull * first, disp, 3ull * count * sizeof(double), cudaMemcpyHostToDevice, s));
    // particle::orientation(q) refreshes the world-frame inertia tensors (core/particle.h:40-44, 140-146)
    if (quat && ctx->dyn_enabled) dynamics_derive_kernel<<<div_up(count, 256), 256, 0, s>>>(ctx->d_quat, ctx->dyn, first, count);
    return PK_OK;
}

// ------------------------------------------------------------------------------------ the stage
int pk_collide_resident(pk_ctx *ctx, pk_step_result *out)
{
    if (!ctx) return PK_E_INVALID;
    PkRange range("pk_collide_resident");
    cudaSetDevice(ctx->cfg.device);
    if (ctx->pairs_in_flight)
    {
        // a previous pk_collide failed between the pair sort and pk_fetch_results: let its copy finish
        // before the pair buffers are written again
        cudaStreamWaitEvent(ctx->stream, ctx->ev_pairs_copied, 0);
        ctx->pairs_in_flight = false;
    }
    PK_TRY(upload_shapes(ctx));
    ctx->have_results = false;
    ctx->fetched = false;
    ctx->launches = 0;
    cudaStream_t s = ctx->stream;
    const uint32_t n = ctx->n_bodies;
    const int mode_query = ctx->cfg.mode == PK_MODE_QUERY;
    if (ctx->alive_dirty)
    {
        uint32_t c = 0;
        for (uint32_t i = 0; i < n; ++i) c += (ctx->h_flags[i] & FLAG_ALIVE) ? 1u : 0u;
        ctx->n_alive = c;
        ctx->alive_dirty = false;
    }
    const uint32_t m = ctx->n_alive;
    WorldTiling wt;
    wt.num_worlds = ctx->cfg.num_worlds;
    wt.grid = 1;
    while (static_cast<uint64_t>(wt.grid) * wt.grid * wt.grid < wt.num_worlds) ++wt.grid;
    const uint32_t *d_world = (ctx->cfg.num_worlds > 1 && ctx->have_world) ? ctx->d_world : nullptr;

    nvtxRangePushA("pk: broadphase (bounds, LBVH, overlap, pair sort)");
    struct PopRange
    {
        bool armed = true;
        ~PopRange()
        {
            if (armed) nvtxRangePop();
        }
    } broad_range;
    cudaEventRecord(ctx->ev[ST_BOUNDS], s);
    scene_reset_kernel<<<1, 64, 0, s>>>(ctx->d_scene, ctx->d_counters, C_COUNT);
    ctx->launches += 1;
    if (n)
    {
        bounds_fat_kernel<<<div_up(n, 256), 256, 0, s>>>(ctx->d_shapes, ctx->d_pos, ctx->d_quat, ctx->d_disp, ctx->d_shape_id,
                                                         ctx->d_flags, n, mode_query, ctx->epoch, ctx->st, ctx->d_scene,
                                                         ctx->d_counters + C_MOVED);
        ctx->launches += 1;
    }
    cudaEventRecord(ctx->ev[ST_MORTON], s);
    ctx->tree_m = 0;
    ctx->tree_valid = true;
    uint64_t npairs = 0; // the pair count, or (speculate) the bound the grids are sized for
    const unsigned long long *pairs_dev = nullptr;
    const bool speculate = ctx->pair_guess_valid && ctx->pair_guess > 0 && !ctx->pairs_read_back && !getenv("PK_SYNC_PAIRS");
    int pair_buf = 0;
    if (m >= 2)
    {
        morton_kernel<<<div_up(n, 256), 256, 0, s>>>(ctx->st.stored, ctx->st.alive, d_world, n, ctx->d_scene, wt,
                                                     ctx->d_bkeys[0], ctx->d_bvals[0]);
        ctx->launches += 1;
        cudaEventRecord(ctx->ev[ST_BODY_SORT], s);
        int bres = 0;
        PK_TRY(radix_sort(ctx, ctx->d_bkeys, ctx->d_bvals, n, {0, 8, 16, 24}, &bres));
        cudaEventRecord(ctx->ev[ST_LEAVES], s);
        leaf_kernel<<<div_up(m, 256), 256, 0, s>>>(ctx->d_bvals[bres], m, ctx->st.stored, ctx->st.last_move, ctx->st.create,
                                                   d_world, ctx->d_scene, wt, ctx->d_leaves, ctx->d_nodes, ctx->d_merge_flag);
        cudaEventRecord(ctx->ev[ST_HIERARCHY], s);
        hierarchy_kernel<<<div_up(m, 256), 256, 0, s>>>(ctx->d_bkeys[bres], m, ctx->d_nodes, ctx->d_right, ctx->d_range_last,
                                                        ctx->d_merge_flag, ctx->d_root);
        rope_kernel<<<div_up(2ull * m - 1, 256), 256, 0, s>>>(m, ctx->d_nodes, ctx->d_right, ctx->d_range_last);
        ctx->tree_m = m;
        cudaEventRecord(ctx->ev[ST_OVERLAP], s);
        uint32_t p_begin = static_cast<uint32_t>(static_cast<uint64_t>(m) * ctx->cfg.shard_rank / ctx->cfg.shard_count);
        uint32_t p_end = static_cast<uint32_t>(static_cast<uint64_t>(m) * (ctx->cfg.shard_rank + 1) / ctx->cfg.shard_count);
        const bool use_rows = !ctx->rows_off;
        if (p_end > p_begin)
        {
            const uint32_t blocks = div_up(p_end - p_begin, OVERLAP_THREADS);
            PairRows pr{ctx->d_row_count, ctx->d_rows, ctx->d_counters + C_ROW_OVERFLOW};
            if (use_rows)
            {
                cudaMemsetAsync(ctx->d_row_count, 0, static_cast<size_t>(n) * sizeof(uint32_t), s);
                overlap_kernel<true><<<blocks, OVERLAP_THREADS, 0, s>>>(ctx->d_nodes, ctx->d_leaves, m, p_begin, p_end, mode_query, ctx->d_pkeys[0],
                                                                        ctx->cfg.max_pairs, ctx->d_counters + C_PAIRS, pr);
            }
            else
                overlap_kernel<false><<<blocks, OVERLAP_THREADS, 0, s>>>(ctx->d_nodes, ctx->d_leaves, m, p_begin, p_end, mode_query, ctx->d_pkeys[0],
                                                                         ctx->cfg.max_pairs, ctx->d_counters + C_PAIRS, pr);
        }
        ctx->launches += 4;
        PK_CUDA(cudaGetLastError());
        cudaEventRecord(ctx->ev[ST_PAIR_SORT], s);
        if (speculate)
        {
            // grids for 1/8 more pairs than last step found (+64 Ki); the kernels read the count on the device
            npairs = std::min<uint64_t>(ctx->cfg.max_pairs, ctx->pair_guess + ctx->pair_guess / 8 + 65536);
            pairs_dev = ctx->d_counters + C_PAIRS;
        }
        else
        {
            PK_TRY(read_counters(ctx));
            npairs = ctx->h_counters[C_PAIRS];
            if (npairs > ctx->cfg.max_pairs) return pair_overflow(ctx, out, npairs);
            if (ctx->h_counters[C_ROW_OVERFLOW]) return repeat_step(ctx, out, true);
        }
        if (npairs && use_rows && p_end > p_begin)
        {
            // the rows one after the other, each sorted: ascending keys (pk_broadphase.cuh, K6 ROWS)
            const uint32_t nt = div_up(n, ROWS_TILE);
            pair_rows_sum_kernel<<<nt, ROWS_TILE, 0, s>>>(ctx->d_row_count, n, ctx->d_row_tiles);
            tile_sum_scan_kernel<<<1, 256, 0, s>>>(ctx->d_row_tiles, nt, nullptr);
            pair_rows_emit_kernel<<<nt, ROWS_TILE, 0, s>>>(ctx->d_row_count, ctx->d_rows, n, ctx->d_row_tiles, ctx->d_pkeys[1], ctx->cfg.max_pairs);
            ctx->launches += 4;
            pair_buf = 1;
            PK_CUDA(cudaGetLastError());
        }
        else if (npairs)
        {
            // (min id, max id) fields packed to 2·idbits significant bits: ceil(2·idbits / 8) passes
            int idbits = bits_for(n);
            std::vector<int> shifts;
            for (int b = 0; b < 2 * idbits; b += 8) shifts.push_back(b);
            PK_TRY(radix_sort(ctx, ctx->d_pkeys, nullptr, npairs, shifts, &pair_buf, idbits, pairs_dev));
        }
    }
    else
    {
        for (int k = ST_BODY_SORT; k <= ST_PAIR_SORT; ++k) cudaEventRecord(ctx->ev[k], s);
    }
    ctx->d_pairs_sorted = ctx->d_pkeys[pair_buf];
    ctx->pairs_in_flight = false;
    nvtxRangePop();
    broad_range.armed = false;
    if (ctx->want_host_results && npairs)
    {
        // pk_collide: the sorted pair keys are final here; ship them to the host on a second stream while
        // the narrowphase (which only reads them) runs (npairs may be the bound of a count still on the device: the
        // keys past the count are never looked at)
        PK_TRY(ensure_host_pairs(ctx, npairs));
        cudaEventRecord(ctx->ev_sorted, s);
        cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_sorted, 0);
        PK_CUDA(cudaMemcpyAsync(ctx->h_pairs, ctx->d_pairs_sorted, npairs * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->copy_stream));
        cudaEventRecord(ctx->ev_pairs_copied, ctx->copy_stream);
        ctx->pairs_in_flight = true;
    }
    // pk_collide: the EPA kernels also deliver every finished contact record to the caller-visible pinned buffer
    // (one coalesced 88-byte store per record, spread over the time the kernels run), so that the device→host
    // copy of the contacts does not have to wait for them.  Only when a pinned buffer for max_contacts records
    // is affordable; a step whose contacts need compaction (a GJK hit without an EPA result) falls back to the copy.
    ContactRec *mirror = nullptr;
    ctx->contacts_mirrored = false;
    if (ctx->want_host_results && npairs && ctx->max_contacts * sizeof(pk_contact) <= (2ull << 30) && !getenv("PK_NO_MIRROR"))
    {
        if (ctx->h_contacts_cap < ctx->max_contacts)
        {
            if (ctx->h_contacts) cudaFreeHost(ctx->h_contacts);
            ctx->h_contacts = nullptr;
            ctx->h_contacts_cap = 0;
            PK_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_contacts), ctx->max_contacts * sizeof(pk_contact), cudaHostAllocDefault));
            ctx->h_contacts_cap = ctx->max_contacts;
        }
        mirror = reinterpret_cast<ContactRec *>(ctx->h_contacts);
        ctx->contacts_mirrored = true;
    }
    PK_TRY(run_narrowphase(ctx, ctx->d_pairs_sorted, nullptr, nullptr, npairs, true, mirror, pairs_dev));
    PK_TRY(read_counters(ctx));
    if (getenv("PK_DEBUG"))
        fprintf(stderr, "[pk] pairs %llu, to gjk_kernel %llu (%llu %llu %llu %llu), hits %llu\n", ctx->h_counters[C_PAIRS],
                ctx->h_counters[C_GJK_CLASS] + ctx->h_counters[C_GJK_CLASS + 1] + ctx->h_counters[C_GJK_CLASS + 2] + ctx->h_counters[C_GJK_CLASS + 3],
                ctx->h_counters[C_GJK_CLASS], ctx->h_counters[C_GJK_CLASS + 1], ctx->h_counters[C_GJK_CLASS + 2], ctx->h_counters[C_GJK_CLASS + 3],
                ctx->h_counters[C_HITS]);
    if (pairs_dev)
    {
        const uint64_t found = ctx->h_counters[C_PAIRS];
        if (found > ctx->cfg.max_pairs) return pair_overflow(ctx, out, found);
        // more pairs than the grids were sized for, or a full row: the same step again with the count read back
        if (found > npairs || ctx->h_counters[C_ROW_OVERFLOW]) return repeat_step(ctx, out, ctx->h_counters[C_ROW_OVERFLOW] != 0);
        npairs = found;
    }
    ctx->pair_guess = npairs;
    ctx->pair_guess_valid = true;
    uint64_t hits = ctx->h_counters[C_HITS];
    uint64_t valid = ctx->h_counters[C_VALID];
    uint64_t over = ctx->h_counters[C_EPA_OVERFLOW];
    int status = PK_OK;
    if (hits > ctx->max_contacts)
    {
        ctx->last_error = "GJK hits exceed pk_config.max_contacts";
        status = PK_E_PAIR_OVERFLOW;
        hits = ctx->max_contacts;
    }
    ctx->d_contacts_final = ctx->d_contacts[0];
    if (status != PK_OK) ctx->contacts_mirrored = false;
    if (status == PK_OK && valid != hits)
    {
        // some GJK hits ended without a value in EPA (degenerate pad / exhausted heap / overflow)
        uint32_t nt = div_up(hits, SCAN_TILE);
        flag_tile_sum_kernel<<<nt, 256, 0, s>>>(ctx->d_valid, hits, nullptr, ctx->d_scan_tiles);
        tile_sum_scan_kernel<<<1, 256, 0, s>>>(ctx->d_scan_tiles, nt, nullptr);
        flag_scan_apply_kernel<<<nt, 256, 0, s>>>(ctx->d_valid, hits, nullptr, ctx->d_scan_tiles, ctx->d_valid_index);
        compact_records_kernel<<<div_up(hits, 256), 256, 0, s>>>(ctx->d_valid, ctx->d_valid_index, hits,
                                                                 reinterpret_cast<const unsigned char *>(ctx->d_contacts[0]),
                                                                 reinterpret_cast<unsigned char *>(ctx->d_contacts[1]),
                                                                 static_cast<int>(sizeof(ContactRec)));
        ctx->launches += 4;
        ctx->d_contacts_final = ctx->d_contacts[1];
        ctx->contacts_mirrored = false;
        PK_CUDA(cudaGetLastError());
    }
    cudaEventRecord(ctx->ev[ST_FETCH], s);
    cudaEventRecord(ctx->ev[ST_COUNT], s);
    PK_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < ST_COUNT; ++k)
    {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev[k], ctx->ev[k + 1]);
        ctx->stage_ms[k] = ms;
    }
    ctx->num_pairs = npairs;
    ctx->num_contacts = (status == PK_OK) ? valid : 0;
    ctx->have_results = status == PK_OK;
    ctx->device_results = ctx->have_results;
    if (out)
    {
        std::memset(out, 0, sizeof(*out));
        out->num_pairs = npairs;
        out->num_contacts = ctx->num_contacts;
        out->num_moved = ctx->h_counters[C_MOVED];
        out->epa_overflow = over;
        out->gjk_hits = ctx->h_counters[C_HITS];
        float bp = 0.f, np_ms = 0.f;
        for (int k = ST_BOUNDS; k <= ST_PAIR_SORT; ++k) bp += ctx->stage_ms[k];
        for (int k = ST_GJK; k <= ST_COMPACT; ++k) np_ms += ctx->stage_ms[k];
        out->ms_broadphase = bp;
        out->ms_narrowphase = np_ms;
        out->ms_total = bp + np_ms;
        out->step_index = static_cast<uint32_t>(ctx->epoch);
    }
    if (status != PK_E_PAIR_OVERFLOW) ctx->epoch += 1; // an overflowed step is repeated after pk_reserve_pairs
    if (status == PK_OK && over)
    {
        ctx->last_error = "EPA polytope exceeded the per-pair scratch for some pairs";
        status = PK_E_EPA_OVERFLOW;
    }
    return status;
}

int pk_fetch_results(pk_ctx *ctx)
{
    if (!ctx) return PK_E_INVALID;
    PkRange range("pk_fetch_results");
    if (!ctx->have_results) return PK_E_STATE;
    if (ctx->fetched) return PK_OK;
    if (!ctx->device_results) return PK_E_STATE;
    cudaSetDevice(ctx->cfg.device);
    cudaEventRecord(ctx->ev[ST_FETCH], ctx->stream);
    PK_TRY(ensure_host_pairs(ctx, ctx->num_pairs));
    if (ctx->num_contacts > ctx->h_contacts_cap)
    {
        if (ctx->h_contacts) cudaFreeHost(ctx->h_contacts);
        ctx->h_contacts = nullptr;
        size_t cap = std::min<size_t>(ctx->max_contacts, std::max<size_t>(ctx->num_contacts * 5 / 4, 1024));
        PK_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_contacts), cap * sizeof(pk_contact), cudaHostAllocDefault));
        ctx->h_contacts_cap = cap;
    }
    if (ctx->pairs_in_flight)
        cudaStreamWaitEvent(ctx->stream, ctx->ev_pairs_copied, 0); // already on their way since the pair sort
    else if (ctx->num_pairs)
        PK_CUDA(cudaMemcpyAsync(ctx->h_pairs, ctx->d_pairs_sorted, ctx->num_pairs * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                                ctx->stream));
    ctx->pairs_in_flight = false;
    if (ctx->num_contacts && !ctx->contacts_mirrored) // (mirrored: the EPA kernels already delivered them)
        PK_CUDA(cudaMemcpyAsync(ctx->h_contacts, ctx->d_contacts_final, ctx->num_contacts * sizeof(pk_contact),
                                cudaMemcpyDeviceToHost, ctx->stream));
    cudaEventRecord(ctx->ev[ST_COUNT], ctx->stream);
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->stage_ms[ST_FETCH], ctx->ev[ST_FETCH], ctx->ev[ST_COUNT]);
    ctx->fetched = true;
    return PK_OK;
}

int pk_collide(pk_ctx *ctx, pk_step_result *out)
{
    if (!ctx) return PK_E_INVALID;
    ctx->want_host_results = true;
    int s = pk_collide_resident(ctx, out);
    ctx->want_host_results = false;
    if (s != PK_OK && s != PK_E_EPA_OVERFLOW) return s;
    int f = pk_fetch_results(ctx);
    return f != PK_OK ? f : s;
}

int pk_pairs(pk_ctx *ctx, const uint64_t **keys, uint64_t *n)
{
    if (!ctx || !keys || !n) return PK_E_INVALID;
    if (!ctx->have_results || !ctx->fetched) return PK_E_STATE;
    *keys = ctx->h_pairs;
    *n = ctx->num_pairs;
    return PK_OK;
}

int pk_contacts(pk_ctx *ctx, const pk_contact **recs, uint64_t *n)
{
    if (!ctx || !recs || !n) return PK_E_INVALID;
    if (!ctx->have_results || !ctx->fetched) return PK_E_STATE;
    *recs = ctx->h_contacts;
    *n = ctx->num_contacts;
    return PK_OK;
}

int pk_pairs_device(pk_ctx *ctx, const void **dptr, uint64_t *n)
{
    if (!ctx || !dptr || !n) return PK_E_INVALID;
    if (!ctx->have_results) return PK_E_STATE;
    *dptr = ctx->d_pairs_sorted; // the pair keys are not touched by pk_gjk_epa_batch
    *n = ctx->num_pairs;
    return PK_OK;
}

int pk_contact_points(pk_ctx *ctx, const pk_contact_point **pts, uint64_t *n)
{
    if (!ctx || !pts || !n) return PK_E_INVALID;
    if (!ctx->have_results || !ctx->device_results) return PK_E_STATE;
    cudaSetDevice(ctx->cfg.device);
    static_assert(sizeof(ContactPointRec) == sizeof(pk_contact_point), "contact point layouts differ");
    const uint64_t m = ctx->num_contacts;
    if (m > ctx->points_cap)
    {
        if (ctx->d_points) cudaFree(ctx->d_points);
        if (ctx->h_points) cudaFreeHost(ctx->h_points);
        ctx->d_points = nullptr;
        ctx->h_points = nullptr;
        ctx->points_cap = 0;
        const size_t cap = std::min<size_t>(ctx->max_contacts, std::max<size_t>(m * 5 / 4, 1024));
        PK_CUDA(cudaMalloc(reinterpret_cast<void **>(&ctx->d_points), cap * sizeof(ContactPointRec)));
        PK_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_points), cap * sizeof(pk_contact_point), cudaHostAllocDefault));
        ctx->points_cap = cap;
    }
    if (m)
    {
        contact_points_kernel<<<div_up(m, 256), 256, 0, ctx->stream>>>(ctx->d_contacts_final, m, ctx->d_pos, ctx->d_quat, ctx->d_points);
        PK_CUDA(cudaGetLastError());
        PK_CUDA(cudaMemcpyAsync(ctx->h_points, ctx->d_points, m * sizeof(pk_contact_point), cudaMemcpyDeviceToHost, ctx->stream));
        PK_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    *pts = ctx->h_points;
    *n = m;
    return PK_OK;
}

// ------------------------------------------------------------------------------------ several GPUs, one process
struct pk_multi
{
    std::vector<pk_ctx *> ctx;
};

int pk_create_multi(const pk_config *cfg, const int *devices, int n, pk_multi **out)
{
    if (!cfg || !devices || n < 1 || !out) return PK_E_INVALID;
    pk_multi *m = new (std::nothrow) pk_multi();
    if (!m) return PK_E_OOM;
    for (int i = 0; i < n; ++i)
    {
        pk_config c = *cfg;
        c.device = devices[i];
        c.shard_rank = static_cast<uint32_t>(i);
        c.shard_count = static_cast<uint32_t>(n);
        pk_ctx *x = nullptr;
        const int s = pk_create(&c, &x);
        if (s != PK_OK)
        {
            for (pk_ctx *y : m->ctx) pk_destroy(y);
            delete m;
            return s;
        }
        m->ctx.push_back(x);
    }
    *out = m;
    return PK_OK;
}

int pk_destroy_multi(pk_multi *m)
{
    if (!m) return PK_E_INVALID;
    int s = PK_OK;
    for (pk_ctx *x : m->ctx)
    {
        const int r = pk_destroy(x);
        if (s == PK_OK) s = r;
    }
    delete m;
    return s;
}

int pk_multi_size(pk_multi *m, int *n)
{
    if (!m || !n) return PK_E_INVALID;
    *n = static_cast<int>(m->ctx.size());
    return PK_OK;
}

int pk_multi_ctx(pk_multi *m, int i, pk_ctx **ctx)
{
    if (!m || !ctx || i < 0 || i >= static_cast<int>(m->ctx.size())) return PK_E_INVALID;
    *ctx = m->ctx[static_cast<size_t>(i)];
    return PK_OK;
}

int pk_multi_bodies_resize(pk_multi *m, uint32_t n)
{
    if (!m) return PK_E_INVALID;
    for (pk_ctx *x : m->ctx) PK_TRY(pk_bodies_resize(x, n));
    return PK_OK;
}

int pk_multi_bodies_upload(pk_multi *m, const double *pos, const double *quat, const double *disp, const uint32_t *shape_id,
                           const uint8_t *flags, const uint32_t *world_id, uint32_t first, uint32_t count)
{
    if (!m) return PK_E_INVALID;
    for (pk_ctx *x : m->ctx) PK_TRY(pk_bodies_upload(x, pos, quat, disp, shape_id, flags, world_id, first, count)); // async copies, one stream per device
    return PK_OK;
}

int pk_multi_bodies_update_pose(pk_multi *m, const double *pos, const double *quat, const double *disp, uint32_t first, uint32_t count)
{
    if (!m) return PK_E_INVALID;
    for (pk_ctx *x : m->ctx) PK_TRY(pk_bodies_update_pose(x, pos, quat, disp, first, count));
    return PK_OK;
}

int pk_multi_collide(pk_multi *m, pk_step_result *total)
{
    if (!m) return PK_E_INVALID;
    const size_t n = m->ctx.size();
    std::vector<pk_step_result> res(n);
    std::vector<int> status(n, PK_OK);
    std::vector<std::thread> th;
    th.reserve(n);
    for (size_t i = 0; i < n; ++i)
    {
        try
        {
            th.emplace_back([&, i] { status[i] = pk_collide(m->ctx[i], &res[i]); });
        }
        catch (...) // no thread to be had: nothing may cross the C boundary, the context is stepped here instead
        {
            status[i] = pk_collide(m->ctx[i], &res[i]);
        }
    }
    for (std::thread &t : th) t.join();
    if (total)
    {
        std::memset(total, 0, sizeof(*total));
        for (size_t i = 0; i < n; ++i)
        {
            total->num_pairs += res[i].num_pairs;
            total->num_contacts += res[i].num_contacts;
            total->gjk_hits += res[i].gjk_hits;
            total->epa_overflow += res[i].epa_overflow;
            total->pairs_required = std::max(total->pairs_required, res[i].pairs_required);
            total->num_moved = std::max(total->num_moved, res[i].num_moved); // every context updates all bodies
            total->ms_broadphase = std::max(total->ms_broadphase, res[i].ms_broadphase);
            total->ms_narrowphase = std::max(total->ms_narrowphase, res[i].ms_narrowphase);
            total->ms_total = std::max(total->ms_total, res[i].ms_total);
            total->step_index = res[i].step_index;
        }
    }
    for (size_t i = 0; i < n; ++i)
        if (status[i] != PK_OK) return status[i];
    return PK_OK;
}

// ------------------------------------------------------------------------------------ integrator
namespace
{
// Eigen Matrix3d::inverse (compute_inverse_size3_helper): cofactors of column 0, det = their dot with column 0,
// every entry = cofactor · (1 / det); particle.h:22-23.  Host side, once per upload, like the reference's
// constructor.  -ffp-contract is off for host code of this file (nvcc passes -fmad only to the device).
void invert3(const double *a, double *r)
{
    auto A = [&](int i, int j) { return a[3 * i + j]; };
    auto cof = [&](int i, int j)
    {
        int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        volatile double p = A(i1, j1) * A(i2, j2), q = A(i1, j2) * A(i2, j1); // two roundings, no contraction
        return p - q;
    };
    volatile double t0 = cof(0, 0) * A(0, 0), t1 = cof(1, 0) * A(1, 0), t2 = cof(2, 0) * A(2, 0);
    const double det = (t0 + t1) + t2;
    const double invdet = 1.0 / det;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r[3 * j + i] = cof(i, j) * invdet;
}
} // namespace

int pk_dynamics_enable(pk_ctx *ctx)
{
    if (!ctx) return PK_E_INVALID;
    if (ctx->dyn_enabled) return PK_OK;
    cudaSetDevice(ctx->cfg.device);
    const size_t nb = ctx->cfg.max_bodies;
    PK_TRY(dev_alloc(ctx, &ctx->dyn.vel, 3 * nb));
    PK_TRY(dev_alloc(ctx, &ctx->dyn.ang_vel, 3 * nb));
    PK_TRY(dev_alloc(ctx, &ctx->dyn.acc, 3 * nb));
    PK_TRY(dev_alloc(ctx, &ctx->dyn.torque, 3 * nb));
    PK_TRY(dev_alloc(ctx, &ctx->dyn.mass, 2 * nb));
    PK_TRY(dev_alloc(ctx, &ctx->dyn.inertia, 18 * nb));
    PK_TRY(dev_alloc(ctx, &ctx->dyn.inertia_w, 18 * nb));
    PK_TRY(dev_alloc(ctx, &ctx->d_material, 2 * nb));
    {
        std::vector<double> half(2 * nb, 0.5); // object_desc defaults: restitution 0.5, friction 0.5 (core/object.h:107-108)
        PK_CUDA(cudaMemcpy(ctx->d_material, half.data(), half.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    for (double *p : {ctx->dyn.vel, ctx->dyn.ang_vel, ctx->dyn.acc, ctx->dyn.torque})
        PK_CUDA(cudaMemsetAsync(p, 0, 3 * nb * sizeof(double), ctx->stream));
    PK_CUDA(cudaEventCreate(&ctx->ev_dyn[0]));
    PK_CUDA(cudaEventCreate(&ctx->ev_dyn[1]));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->dyn_enabled = true;
    return PK_OK;
}

int pk_dynamics_upload(pk_ctx *ctx, const double *vel, const double *ang_vel, const double *mass, const double *inertia_local,
                       uint32_t first, uint32_t count)
{
    if (!ctx || !vel || !ang_vel || !mass || !inertia_local) return PK_E_INVALID;
    if (!ctx->dyn_enabled) return PK_E_STATE;
    if (static_cast<uint64_t>(first) + count > ctx->n_bodies) return PK_E_INVALID;
    if (count == 0) return PK_OK;
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t s = ctx->stream;
    std::vector<double> mm(2ull * count), in(18ull * count);
    for (uint32_t i = 0; i < count; ++i)
    {
        mm[2ull * i] = mass[i];
        mm[2ull * i + 1] = 1.0 / mass[i]; // particle.h:21
        std::memcpy(&in[18ull * i], inertia_local + 9ull * i, 9 * sizeof(double));
        if (mm[2ull * i + 1] == 0.0)
            std::fill(&in[18ull * i + 9], &in[18ull * i + 18], 0.0); // infinite mass: zero inverse tensor (particle.h:25-26)
        else
            invert3(inertia_local + 9ull * i, &in[18ull * i + 9]);
    }
    PK_CUDA(cudaMemcpyAsync(ctx->dyn.vel + 3ull * first, vel, 3ull * count * sizeof(double), cudaMemcpyHostToDevice, s));
    PK_CUDA(cudaMemcpyAsync(ctx->dyn.ang_vel + 3ull * first, ang_vel, 3ull * count * sizeof(double), cudaMemcpyHostToDevice, s));
    PK_CUDA(cudaMemcpyAsync(ctx->dyn.mass + 2ull * first, mm.data(), mm.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    PK_CUDA(cudaMemcpyAsync(ctx->dyn.inertia + 18ull * first, in.data(), in.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    // update_derived_state from the poses already uploaded (particle.h:27): call after pk_bodies_upload
    dynamics_derive_kernel<<<div_up(count, 256), 256, 0, s>>>(ctx->d_quat, ctx->dyn, first, count);
    PK_CUDA(cudaGetLastError());
    PK_CUDA(cudaStreamSynchronize(s)); // mm / in are stack-owned
    return PK_OK;
}

int pk_dynamics_set_velocities(pk_ctx *ctx, const double *vel, const double *ang_vel, uint32_t first, uint32_t count)
{
    if (!ctx) return PK_E_INVALID;
    if (!ctx->dyn_enabled) return PK_E_STATE;
    if (static_cast<uint64_t>(first) + count > ctx->n_bodies) return PK_E_INVALID;
    if (count == 0) return PK_OK;
    cudaSetDevice(ctx->cfg.device);
    if (vel) PK_CUDA(cudaMemcpyAsync(ctx->dyn.vel + 3ull * first, vel, 3ull * count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (ang_vel)
        PK_CUDA(cudaMemcpyAsync(ctx->dyn.ang_vel + 3ull * first, ang_vel, 3ull * count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return PK_OK;
}

int pk_dynamics_set_forces(pk_ctx *ctx, const double *acc, const double *torque, uint32_t first, uint32_t count)
{
    if (!ctx) return PK_E_INVALID;
    if (!ctx->dyn_enabled) return PK_E_STATE;
    if (static_cast<uint64_t>(first) + count > ctx->n_bodies) return PK_E_INVALID;
    if (count == 0) return PK_OK;
    cudaSetDevice(ctx->cfg.device);
    if (acc) PK_CUDA(cudaMemcpyAsync(ctx->dyn.acc + 3ull * first, acc, 3ull * count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (torque)
        PK_CUDA(cudaMemcpyAsync(ctx->dyn.torque + 3ull * first, torque, 3ull * count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return PK_OK;
}

int pk_integrate_velocities(pk_ctx *ctx, double dt, const double gravity[3])
{
    if (!ctx || !gravity) return PK_E_INVALID;
    if (!ctx->dyn_enabled) return PK_E_STATE;
    cudaSetDevice(ctx->cfg.device);
    const uint32_t n = ctx->n_bodies;
    cudaEventRecord(ctx->ev_dyn[0], ctx->stream);
    if (n)
    {
        integrate_vel_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(ctx->d_flags, ctx->dyn, ctx->d_disp, n, dt,
                                                                     d3{gravity[0], gravity[1], gravity[2]});
        PK_CUDA(cudaGetLastError());
    }
    cudaEventRecord(ctx->ev_dyn[1], ctx->stream);
    return PK_OK;
}

int pk_integrate_positions(pk_ctx *ctx, double dt)
{
    if (!ctx) return PK_E_INVALID;
    if (!ctx->dyn_enabled) return PK_E_STATE;
    cudaSetDevice(ctx->cfg.device);
    const uint32_t n = ctx->n_bodies;
    cudaEventRecord(ctx->ev_dyn[0], ctx->stream);
    if (n)
    {
        integrate_pos_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(ctx->d_flags, ctx->dyn, ctx->d_pos, ctx->d_quat, n, dt);
        PK_CUDA(cudaGetLastError());
    }
    cudaEventRecord(ctx->ev_dyn[1], ctx->stream);
    return PK_OK;
}

int pk_dynamics_download(pk_ctx *ctx, double *pos, double *quat, double *vel, double *ang_vel, uint32_t first, uint32_t count)
{
    if (!ctx) return PK_E_INVALID;
    if (static_cast<uint64_t>(first) + count > ctx->n_bodies) return PK_E_INVALID;
    if ((vel || ang_vel) && !ctx->dyn_enabled) return PK_E_STATE;
    if (count == 0) return PK_OK;
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t s = ctx->stream;
    if (pos) PK_CUDA(cudaMemcpyAsync(pos, ctx->d_pos + 3ull * first, 3ull * count * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (quat) PK_CUDA(cudaMemcpyAsync(quat, ctx->d_quat + 4ull * first, 4ull * count * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (vel) PK_CUDA(cudaMemcpyAsync(vel, ctx->dyn.vel + 3ull * first, 3ull * count * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (ang_vel)
        PK_CUDA(cudaMemcpyAsync(ang_vel, ctx->dyn.ang_vel + 3ull * first, 3ull * count * sizeof(double), cudaMemcpyDeviceToHost, s));
    PK_CUDA(cudaStreamSynchronize(s));
    if (ctx->dyn_enabled && ctx->ev_dyn[0]) cudaEventElapsedTime(&ctx->dyn_ms, ctx->ev_dyn[0], ctx->ev_dyn[1]);
    return PK_OK;
}

int pk_displacements(pk_ctx *ctx, double *disp, uint32_t first, uint32_t count)
{
    if (!ctx || !disp) return PK_E_INVALID;
    if (static_cast<uint64_t>(first) + count > ctx->n_bodies) return PK_E_INVALID;
    if (count == 0) return PK_OK;
    cudaSetDevice(ctx->cfg.device);
    PK_CUDA(cudaMemcpyAsync(disp, ctx->d_disp + 3ull * first, 3ull * count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    return PK_OK;
}

// ------------------------------------------------------------------------------------ contact rows
int pk_material_upload(pk_ctx *ctx, const double *restitution, const double *friction, uint32_t first, uint32_t count)
{
    if (!ctx || !restitution || !friction) return PK_E_INVALID;
    if (!ctx->dyn_enabled) return PK_E_STATE;
    if (static_cast<uint64_t>(first) + count > ctx->n_bodies) return PK_E_INVALID;
    if (count == 0) return PK_OK;
    cudaSetDevice(ctx->cfg.device);
    std::vector<double> m(2ull * count);
    for (uint32_t i = 0; i < count; ++i)
    {
        m[2ull * i] = restitution[i];
        m[2ull * i + 1] = friction[i];
    }
    PK_CUDA(cudaMemcpyAsync(ctx->d_material + 2ull * first, m.data(), m.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    return PK_OK;
}

int pk_contact_rows_setup(pk_ctx *ctx, double dt, double gravity_norm, uint64_t *nrows)
{
    if (!ctx) return PK_E_INVALID;
    if (!ctx->dyn_enabled || !ctx->man_cap)
    {
        ctx->last_error = "pk_contact_rows_setup needs pk_dynamics_enable and pk_manifolds_enable";
        return PK_E_STATE;
    }
    static_assert(sizeof(SolverPoint) == sizeof(pk_solver_point), "solver point layouts differ");
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t s = ctx->stream;
    if (!ctx->d_sol_rows)
    {
        const uint64_t cap = 4 * ctx->man_cap;
        PK_TRY(dev_alloc(ctx, &ctx->d_sol_valid, cap + 16));
        PK_TRY(dev_alloc(ctx, &ctx->d_sol_index, cap));
        PK_TRY(dev_alloc(ctx, &ctx->d_sol_slot, cap));
        PK_TRY(dev_alloc(ctx, &ctx->d_sol_tiles, div_up(cap, SCAN_TILE) + 1));
        PK_TRY(dev_alloc(ctx, &ctx->d_sol_total, 1));
        PK_TRY(dev_alloc(ctx, &ctx->d_sol_rows, cap));
        PK_CUDA(cudaEventCreate(&ctx->ev_sol[0]));
        PK_CUDA(cudaEventCreate(&ctx->ev_sol[1]));
        ctx->sol_cap = cap;
    }
    const uint64_t nman = ctx->man_count, slots = 4 * nman;
    ctx->sol_count = 0;
    cudaEventRecord(ctx->ev_sol[0], s);
    if (slots)
    {
        const ManifoldRec *man = ctx->d_man[ctx->man_cur];
        const uint32_t nt = div_up(slots, SCAN_TILE);
        solver_valid_kernel<<<div_up(slots, 256), 256, 0, s>>>(man, nman, ctx->d_pos, ctx->d_quat, ctx->d_sol_valid);
        flag_tile_sum_kernel<<<nt, 256, 0, s>>>(ctx->d_sol_valid, slots, nullptr, ctx->d_sol_tiles);
        tile_sum_scan_kernel<<<1, 256, 0, s>>>(ctx->d_sol_tiles, nt, ctx->d_sol_total);
        flag_scan_apply_kernel<<<nt, 256, 0, s>>>(ctx->d_sol_valid, slots, nullptr, ctx->d_sol_tiles, ctx->d_sol_index);
        solver_compact_kernel<<<div_up(slots, 256), 256, 0, s>>>(ctx->d_sol_valid, ctx->d_sol_index, slots, ctx->d_sol_slot);
        // one thread per row; the grid covers the point count of the manifolds (≥ rows), surplus threads leave at once
        solver_rows_kernel<<<div_up(slots, 128), 128, 0, s>>>(man, ctx->d_sol_total, ctx->d_pos, ctx->d_quat, ctx->dyn, ctx->d_material,
                                                             ctx->d_sol_slot, dt, gravity_norm, ctx->d_sol_rows);
        ctx->launches += 6;
        PK_CUDA(cudaGetLastError());
        unsigned long long total = 0;
        cudaEventRecord(ctx->ev_sol[1], s);
        PK_CUDA(cudaMemcpyAsync(&total, ctx->d_sol_total, sizeof(total), cudaMemcpyDeviceToHost, s));
        PK_CUDA(cudaStreamSynchronize(s));
        ctx->sol_count = total;
        cudaEventElapsedTime(&ctx->sol_ms, ctx->ev_sol[0], ctx->ev_sol[1]);
    }
    if (nrows) *nrows = ctx->sol_count;
    return PK_OK;
}

int pk_contact_rows(pk_ctx *ctx, const pk_solver_point **rows, uint64_t *n)
{
    if (!ctx || !rows || !n) return PK_E_INVALID;
    if (!ctx->d_sol_rows) return PK_E_STATE;
    cudaSetDevice(ctx->cfg.device);
    const uint64_t m = ctx->sol_count;
    if (m > ctx->h_sol_cap)
    {
        if (ctx->h_sol_rows) cudaFreeHost(ctx->h_sol_rows);
        ctx->h_sol_rows = nullptr;
        ctx->h_sol_cap = 0;
        const size_t cap = std::min<size_t>(ctx->sol_cap, std::max<size_t>(m * 5 / 4, 1024));
        PK_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_sol_rows), cap * sizeof(pk_solver_point), cudaHostAllocDefault));
        ctx->h_sol_cap = cap;
    }
    if (m)
    {
        PK_CUDA(cudaMemcpyAsync(ctx->h_sol_rows, ctx->d_sol_rows, m * sizeof(pk_solver_point), cudaMemcpyDeviceToHost, ctx->stream));
        PK_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    *rows = ctx->h_sol_rows;
    *n = m;
    return PK_OK;
}

int pk_contact_rows_device(pk_ctx *ctx, const void **dptr, uint64_t *n, float *device_ms)
{
    if (!ctx || !dptr || !n) return PK_E_INVALID;
    if (!ctx->d_sol_rows) return PK_E_STATE;
    *dptr = ctx->d_sol_rows;
    *n = ctx->sol_count;
    if (device_ms) *device_ms = ctx->sol_ms;
    return PK_OK;
}

// ------------------------------------------------------------------------------------ ray casts
int pk_raycast(pk_ctx *ctx, const double *origins, const double *directions, const double *max_distance, const uint32_t *world,
               uint32_t nrays, int mode, pk_ray_hit *hits, uint64_t capacity, uint64_t *nhits)
{
    if (!ctx || !nhits || (nrays && (!origins || !directions || !max_distance)) || (capacity && !hits)) return PK_E_INVALID;
    if (mode != PK_RAY_ALL && mode != PK_RAY_CLOSEST) return PK_E_INVALID;
    if (!ctx->tree_valid)
    {
        ctx->last_error = "pk_raycast needs the tree of a step: call pk_collide / pk_collide_resident first";
        return PK_E_STATE;
    }
    static_assert(sizeof(RayHitRec) == sizeof(pk_ray_hit), "ray hit layouts differ");
    cudaSetDevice(ctx->cfg.device);
    *nhits = 0;
    if (nrays == 0) return PK_OK;
    if (static_cast<size_t>(div_up(std::max<uint64_t>(capacity, 1), SORT_TILE)) * 256 > ctx->tile_hist_entries)
    {
        ctx->last_error = "pk_raycast: capacity exceeds max(max_bodies, max_pairs) of the context";
        return PK_E_INVALID;
    }
    cudaStream_t s = ctx->stream;
    if (!ctx->ev_ray[0])
    {
        PK_CUDA(cudaEventCreate(&ctx->ev_ray[0]));
        PK_CUDA(cudaEventCreate(&ctx->ev_ray[1]));
        PK_TRY(dev_alloc(ctx, &ctx->d_ray_counter, 1));
    }
    if (nrays > ctx->ray_cap)
    {
        for (void *p : {static_cast<void *>(ctx->d_ray_o), static_cast<void *>(ctx->d_ray_d), static_cast<void *>(ctx->d_ray_max),
                        static_cast<void *>(ctx->d_ray_world)})
            if (p) cudaFree(p);
        ctx->d_ray_o = ctx->d_ray_d = ctx->d_ray_max = nullptr;
        ctx->d_ray_world = nullptr;
        ctx->ray_cap = 0;
        const size_t cap = std::max<size_t>(static_cast<size_t>(nrays) * 5 / 4, 1024);
        PK_TRY(dev_alloc(ctx, &ctx->d_ray_o, 3 * cap));
        PK_TRY(dev_alloc(ctx, &ctx->d_ray_d, 3 * cap));
        PK_TRY(dev_alloc(ctx, &ctx->d_ray_max, cap));
        PK_TRY(dev_alloc(ctx, &ctx->d_ray_world, cap));
        ctx->ray_cap = cap;
    }
    if (capacity > ctx->ray_hit_cap)
    {
        for (void *p : {static_cast<void *>(ctx->d_ray_keys[0]), static_cast<void *>(ctx->d_ray_keys[1]), static_cast<void *>(ctx->d_ray_src[0]),
                        static_cast<void *>(ctx->d_ray_src[1]), static_cast<void *>(ctx->d_ray_dist), static_cast<void *>(ctx->d_ray_hits)})
            if (p) cudaFree(p);
        ctx->d_ray_keys[0] = ctx->d_ray_keys[1] = nullptr;
        ctx->d_ray_src[0] = ctx->d_ray_src[1] = nullptr;
        ctx->d_ray_dist = nullptr;
        ctx->d_ray_hits = nullptr;
        ctx->ray_hit_cap = 0;
        PK_TRY(dev_alloc(ctx, &ctx->d_ray_keys[0], capacity));
        PK_TRY(dev_alloc(ctx, &ctx->d_ray_keys[1], capacity));
        PK_TRY(dev_alloc(ctx, &ctx->d_ray_src[0], capacity));
        PK_TRY(dev_alloc(ctx, &ctx->d_ray_src[1], capacity));
        PK_TRY(dev_alloc(ctx, &ctx->d_ray_dist, capacity));
        PK_TRY(dev_alloc(ctx, &ctx->d_ray_hits, capacity));
        ctx->ray_hit_cap = capacity;
    }
    PK_CUDA(cudaMemcpyAsync(ctx->d_ray_o, origins, 3ull * nrays * sizeof(double), cudaMemcpyHostToDevice, s));
    PK_CUDA(cudaMemcpyAsync(ctx->d_ray_d, directions, 3ull * nrays * sizeof(double), cudaMemcpyHostToDevice, s));
    PK_CUDA(cudaMemcpyAsync(ctx->d_ray_max, max_distance, static_cast<size_t>(nrays) * sizeof(double), cudaMemcpyHostToDevice, s));
    const bool worlds = ctx->cfg.num_worlds > 1 && ctx->have_world;
    if (worlds && world)
        PK_CUDA(cudaMemcpyAsync(ctx->d_ray_world, world, static_cast<size_t>(nrays) * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    else if (worlds)
        PK_CUDA(cudaMemsetAsync(ctx->d_ray_world, 0, static_cast<size_t>(nrays) * sizeof(uint32_t), s));
    PK_CUDA(cudaMemsetAsync(ctx->d_ray_counter, 0, sizeof(unsigned long long), s));
    WorldTiling wt;
    wt.num_worlds = worlds ? ctx->cfg.num_worlds : 1;
    wt.grid = 1;
    while (static_cast<uint64_t>(wt.grid) * wt.grid * wt.grid < wt.num_worlds) ++wt.grid;
    cudaEventRecord(ctx->ev_ray[0], s);
    if (ctx->tree_m >= 2)
        ray_cast_kernel<<<div_up(nrays, RAY_THREADS), RAY_THREADS, 0, s>>>(ctx->d_nodes, ctx->d_leaves, ctx->tree_m, ctx->d_root, ctx->d_ray_o,
                                                                          ctx->d_ray_d, ctx->d_ray_max, worlds ? ctx->d_ray_world : nullptr,
                                                                          nrays, mode, wt, ctx->d_scene, ctx->d_ray_keys[0], ctx->d_ray_dist,
                                                                          ctx->d_ray_src[0], capacity, ctx->d_ray_counter);
    else
        ray_brute_kernel<<<div_up(nrays, RAY_THREADS), RAY_THREADS, 0, s>>>(ctx->st.stored, ctx->st.alive, worlds ? ctx->d_world : nullptr,
                                                                           ctx->n_bodies, ctx->d_ray_o, ctx->d_ray_d, ctx->d_ray_max,
                                                                           worlds ? ctx->d_ray_world : nullptr, nrays, mode, ctx->d_ray_keys[0],
                                                                           ctx->d_ray_dist, ctx->d_ray_src[0], capacity, ctx->d_ray_counter);
    ctx->launches += 1;
    PK_CUDA(cudaGetLastError());
    unsigned long long found = 0;
    PK_CUDA(cudaMemcpyAsync(&found, ctx->d_ray_counter, sizeof(found), cudaMemcpyDeviceToHost, s));
    PK_CUDA(cudaStreamSynchronize(s));
    *nhits = found;
    if (found > capacity)
    {
        ctx->last_error = "ray hits exceed the capacity passed to pk_raycast";
        return PK_E_PAIR_OVERFLOW; // *nhits = required capacity
    }
    if (found)
    {
        int buf = 0;
        if (found > 1)
        {
            std::vector<int> shifts;
            const int idbits = bits_for(std::max<uint32_t>(ctx->n_bodies, 2)), raybits = bits_for(std::max<uint32_t>(nrays, 2));
            for (int b = 0; b < idbits; b += 8) shifts.push_back(b);
            for (int b = 0; b < raybits; b += 8) shifts.push_back(32 + b);
            PK_TRY(radix_sort(ctx, ctx->d_ray_keys, ctx->d_ray_src, found, shifts, &buf));
        }
        ray_gather_kernel<<<div_up(found, 256), 256, 0, s>>>(ctx->d_ray_keys[buf], ctx->d_ray_src[buf], ctx->d_ray_dist, found, ctx->d_ray_hits);
        ctx->launches += 1;
        PK_CUDA(cudaGetLastError());
    }
    cudaEventRecord(ctx->ev_ray[1], s);
    if (found) PK_CUDA(cudaMemcpyAsync(hits, ctx->d_ray_hits, found * sizeof(pk_ray_hit), cudaMemcpyDeviceToHost, s));
    PK_CUDA(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&ctx->ray_ms, ctx->ev_ray[0], ctx->ev_ray[1]);
    return PK_OK;
}

int pk_raycast_device_ms(pk_ctx *ctx, float *ms)
{
    if (!ctx || !ms) return PK_E_INVALID;
    *ms = ctx->ray_ms;
    return PK_OK;
}

// ------------------------------------------------------------------------------------ manifolds
int pk_manifolds_enable(pk_ctx *ctx, uint64_t capacity)
{
    // A manifold lives as long as its pair stays in the pair set (collision_phases.h:309-318), but a sharded context
    // only sees the pairs whose lower leaf falls into its slice of the Morton order, and bodies move between slices:
    // manifolds need the gathered pair and contact set of the whole world.
    if (ctx && ctx->cfg.shard_count > 1)
    {
        ctx->last_error = "manifolds need an unsharded context (pk_config.shard_count == 1)";
        return PK_E_STATE;
    }
    if (!ctx || capacity == 0) return PK_E_INVALID;
    if (ctx->man_cap) return PK_E_STATE;
    if (capacity > std::max<uint64_t>(ctx->cfg.max_pairs, ctx->cfg.max_bodies))
    {
        ctx->last_error = "manifold capacity exceeds pk_config.max_pairs (the sort scratch is sized by it)";
        return PK_E_INVALID;
    }
    cudaSetDevice(ctx->cfg.device);
    int s;
#define A(ptr, count)                                   \
    if ((s = dev_alloc(ctx, &(ptr), (count))) != PK_OK) \
    return s
    A(ctx->d_man[0], capacity);
    A(ctx->d_man[1], capacity);
    A(ctx->d_man_stage, 2 * capacity);
    for (int b = 0; b < 2; ++b)
    {
        A(ctx->d_man_ckeys[b], capacity);
        A(ctx->d_man_csrc[b], capacity);
        A(ctx->d_man_began[b], capacity);
        A(ctx->d_man_ended[b], capacity);
    }
    A(ctx->d_man_consumed, ctx->max_contacts + 16);
    A(ctx->d_man_counters, 4);
    A(ctx->d_man_imp, capacity * 12);
#undef A
    ctx->man_cap = capacity;
    ctx->man_count = 0;
    ctx->man_epoch = -1;
    return PK_OK;
}

// narrow_phase::calculate's manifold part (collision_phases.h:252-318) for the step pk_collide* just computed.
int pk_manifolds_update(pk_ctx *ctx, pk_manifold_result *out)
{
    if (!ctx) return PK_E_INVALID;
    if (!ctx->man_cap || !ctx->have_results || !ctx->device_results || ctx->man_epoch == ctx->epoch) return PK_E_STATE;
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t s = ctx->stream;
    const uint64_t m_prev = ctx->man_count;
    const uint64_t nslots = std::min<uint64_t>(ctx->h_counters[C_HITS], ctx->max_contacts);
    const ManifoldRec *prev = ctx->d_man[ctx->man_cur];
    ManifoldRec *next = ctx->d_man[ctx->man_cur ^ 1];
    cudaEventRecord(ctx->ev[0], s);
    PK_CUDA(cudaMemsetAsync(ctx->d_man_counters, 0, 4 * sizeof(unsigned long long), s));
    if (nslots) PK_CUDA(cudaMemsetAsync(ctx->d_man_consumed, 0, nslots, s));
    if (m_prev)
        manifold_old_kernel<<<div_up(m_prev, 128), 128, 0, s>>>(prev, m_prev, ctx->d_pairs_sorted, ctx->num_pairs, ctx->d_hit,
                                                               ctx->d_out_index, ctx->d_valid, ctx->d_contacts[0], ctx->d_pos,
                                                               ctx->d_quat, ctx->d_man_stage, ctx->d_man_consumed,
                                                               ctx->d_man_ckeys[0], ctx->d_man_csrc[0], ctx->man_cap, ctx->d_man_ended[0],
                                                               ctx->d_man_counters);
    if (nslots)
        manifold_new_kernel<<<div_up(nslots, 128), 128, 0, s>>>(ctx->d_contacts[0], ctx->d_valid, ctx->d_man_consumed, nslots, ctx->d_pos,
                                                               ctx->d_quat, ctx->d_man_stage, m_prev, 2 * ctx->man_cap,
                                                               ctx->d_man_ckeys[0], ctx->d_man_csrc[0], ctx->man_cap, ctx->d_man_began[0],
                                                               ctx->d_man_counters);
    PK_CUDA(cudaGetLastError());
    unsigned long long cnt[4] = {0, 0, 0, 0};
    PK_CUDA(cudaMemcpyAsync(cnt, ctx->d_man_counters, sizeof(cnt), cudaMemcpyDeviceToHost, s));
    PK_CUDA(cudaStreamSynchronize(s));
    // the candidate list has man_cap entries: survivors of the old array + new manifolds must fit
    if (cnt[0] > ctx->man_cap || cnt[2] > ctx->man_cap || m_prev + cnt[1] > 2 * ctx->man_cap)
    {
        ctx->last_error = "manifold capacity exceeded (pk_manifolds_enable)";
        return PK_E_PAIR_OVERFLOW;
    }
    const int idbits = bits_for(std::max<uint32_t>(ctx->n_bodies, 2));
    std::vector<int> shifts;
    for (int b = 0; b < 2 * idbits; b += 8) shifts.push_back(b);
    int buf = 0;
    if (cnt[0] > 1) PK_TRY(radix_sort(ctx, ctx->d_man_ckeys, ctx->d_man_csrc, cnt[0], shifts, &buf, idbits));
    if (cnt[0])
        manifold_gather_kernel<<<div_up(cnt[0], 128), 128, 0, s>>>(ctx->d_man_csrc[buf], cnt[0], ctx->d_man_stage, next);
    ctx->man_began_buf = ctx->man_ended_buf = 0;
    if (cnt[2] > 1) PK_TRY(radix_sort(ctx, ctx->d_man_began, nullptr, cnt[2], shifts, &ctx->man_began_buf, idbits));
    if (cnt[3] > 1) PK_TRY(radix_sort(ctx, ctx->d_man_ended, nullptr, cnt[3], shifts, &ctx->man_ended_buf, idbits));
    PK_CUDA(cudaGetLastError());
    cudaEventRecord(ctx->ev[1], s);
    PK_CUDA(cudaStreamSynchronize(s));
    ctx->man_cur ^= 1;
    ctx->man_count = cnt[0];
    ctx->man_began = cnt[2];
    ctx->man_ended = cnt[3];
    ctx->man_epoch = ctx->epoch;
    if (out)
    {
        out->num_manifolds = cnt[0];
        out->num_began = cnt[2];
        out->num_ended = cnt[3];
        cudaEventElapsedTime(&out->ms, ctx->ev[0], ctx->ev[1]);
    }
    return PK_OK;
}

int pk_manifolds(pk_ctx *ctx, const pk_manifold **recs, uint64_t *n)
{
    if (!ctx || !recs || !n) return PK_E_INVALID;
    if (!ctx->man_cap) return PK_E_STATE;
    cudaSetDevice(ctx->cfg.device);
    static_assert(sizeof(pk_manifold) == sizeof(ManifoldRec), "manifold layouts differ");
    const uint64_t m = ctx->man_count;
    if (m > ctx->h_man_cap)
    {
        if (ctx->h_man) cudaFreeHost(ctx->h_man);
        ctx->h_man = nullptr;
        ctx->h_man_cap = 0;
        const size_t cap = std::min<size_t>(ctx->man_cap, std::max<size_t>(m * 5 / 4, 256));
        PK_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_man), cap * sizeof(pk_manifold), cudaHostAllocDefault));
        ctx->h_man_cap = cap;
    }
    if (m)
    {
        PK_CUDA(cudaMemcpyAsync(ctx->h_man, ctx->d_man[ctx->man_cur], m * sizeof(pk_manifold), cudaMemcpyDeviceToHost, ctx->stream));
        PK_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    *recs = ctx->h_man;
    *n = m;
    return PK_OK;
}

int pk_manifolds_device(pk_ctx *ctx, const void **dptr, uint64_t *n)
{
    if (!ctx || !dptr || !n) return PK_E_INVALID;
    if (!ctx->man_cap) return PK_E_STATE;
    *dptr = ctx->d_man[ctx->man_cur];
    *n = ctx->man_count;
    return PK_OK;
}

int pk_manifold_events(pk_ctx *ctx, const uint64_t **began, uint64_t *num_began, const uint64_t **ended, uint64_t *num_ended)
{
    if (!ctx || !began || !num_began || !ended || !num_ended) return PK_E_INVALID;
    if (!ctx->man_cap) return PK_E_STATE;
    cudaSetDevice(ctx->cfg.device);
    const uint64_t nb = ctx->man_began, ne = ctx->man_ended;
    if (nb + ne > ctx->h_man_events_cap)
    {
        if (ctx->h_man_events) cudaFreeHost(ctx->h_man_events);
        ctx->h_man_events = nullptr;
        ctx->h_man_events_cap = 0;
        const size_t cap = std::max<size_t>((nb + ne) * 5 / 4, 256);
        PK_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_man_events), cap * sizeof(uint64_t), cudaHostAllocDefault));
        ctx->h_man_events_cap = cap;
    }
    if (nb)
        PK_CUDA(cudaMemcpyAsync(ctx->h_man_events, ctx->d_man_began[ctx->man_began_buf], nb * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                                ctx->stream));
    if (ne)
        PK_CUDA(cudaMemcpyAsync(ctx->h_man_events + nb, ctx->d_man_ended[ctx->man_ended_buf], ne * sizeof(uint64_t),
                                cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    *began = ctx->h_man_events;
    *num_began = nb;
    *ended = ctx->h_man_events ? ctx->h_man_events + nb : nullptr;
    *num_ended = ne;
    return PK_OK;
}

int pk_manifolds_set_impulses(pk_ctx *ctx, const double *impulses, uint64_t n)
{
    if (!ctx || (!impulses && n)) return PK_E_INVALID;
    if (!ctx->man_cap || n != ctx->man_count) return PK_E_STATE;
    if (!n) return PK_OK;
    cudaSetDevice(ctx->cfg.device);
    PK_CUDA(cudaMemcpyAsync(ctx->d_man_imp, impulses, n * 12 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    manifold_impulses_kernel<<<div_up(4 * n, 256), 256, 0, ctx->stream>>>(ctx->d_man[ctx->man_cur], n, ctx->d_man_imp);
    PK_CUDA(cudaGetLastError());
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    return PK_OK;
}

int pk_contacts_device(pk_ctx *ctx, const void **dptr, uint64_t *n)
{
    if (!ctx || !dptr || !n) return PK_E_INVALID;
    if (!ctx->have_results || !ctx->device_results) return PK_E_STATE;
    *dptr = ctx->d_contacts_final;
    *n = ctx->num_contacts;
    return PK_OK;
}

int pk_stored_bounds(pk_ctx *ctx, double *out6, uint32_t first, uint32_t count)
{
    if (!ctx || !out6) return PK_E_INVALID;
    if (static_cast<uint64_t>(first) + count > ctx->n_bodies) return PK_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    PK_CUDA(cudaMemcpyAsync(out6, ctx->st.stored + 6ull * first, 6ull * count * sizeof(double), cudaMemcpyDeviceToHost,
                            ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    return PK_OK;
}

int pk_stage_times_get(pk_ctx *ctx, pk_stage_times *out)
{
    if (!ctx || !out) return PK_E_INVALID;
    for (int k = 0; k < ST_COUNT; ++k)
    {
        out->ms[k] = ctx->stage_ms[k];
        out->name[k] = kStageNames[k];
    }
    out->launches = ctx->launches;
    out->epa_fallback = static_cast<uint32_t>(ctx->h_counters ? ctx->h_counters[C_EPA_FALLBACK + 1] : 0);
#ifdef PK_ES_REASONS
    if (ctx->h_counters)
        fprintf(stderr, "[epa fallback] to epa_kernel %llu; handed back by either instance: pad %llu tie %llu capacity %llu improper %llu; SCAN iterations thrown away %llu\n",
                static_cast<unsigned long long>(out->epa_fallback), ctx->h_counters[C_EPA_REASONS], ctx->h_counters[C_EPA_REASONS + 1],
                ctx->h_counters[C_EPA_REASONS + 2], ctx->h_counters[C_EPA_REASONS + 3], ctx->h_counters[C_EPA_REASONS + 4]);
#endif
    return PK_OK;
}

int pk_selftest_division(pk_ctx *ctx, uint64_t seed, uint64_t samples, uint64_t *mismatches)
{
    if (!ctx || !mismatches) return PK_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    const uint32_t threads = 256, blocks = static_cast<uint32_t>(ctx->sm_count) * 16;
    const uint32_t per_thread = static_cast<uint32_t>(std::max<uint64_t>(1, samples / (static_cast<uint64_t>(threads) * blocks)));
    PK_CUDA(cudaMemsetAsync(ctx->d_counters + C_COUNT - 1, 0, sizeof(unsigned long long), ctx->stream));
    division_selftest_kernel<<<blocks, threads, 0, ctx->stream>>>(seed, per_thread, ctx->d_counters + C_COUNT - 1);
    unsigned long long h = 0;
    PK_CUDA(cudaMemcpyAsync(&h, ctx->d_counters + C_COUNT - 1, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    *mismatches = h;
    return PK_OK;
}

int pk_stream(pk_ctx *ctx, void **stream)
{
    if (!ctx || !stream) return PK_E_INVALID;
    *stream = ctx->stream;
    return PK_OK;
}

// ------------------------------------------------------------------------------------ one world over several processes
// every rank holds all bodies and rebuilds the tree, traverses its own slice of the sorted leaves and
// runs GJK / EPA on its own pairs; what the ranks exchange is (1) the poses, each rank uploading 1/N of them from the
// host and all-gathering the rest over NVLink instead of N copies of the whole array crossing PCIe, and (2) the
// contact records, which constraint_solver::setup_contacts reads for the whole world (collision/constraint.h:1052-1104).

#define PK_NCCL(call)                                                                            \
    do                                                                                           \
    {                                                                                            \
        ncclResult_t r__ = (call);                                                               \
        if (r__ != ncclSuccess)                                                                  \
        {                                                                                        \
            ctx->last_error = std::string(#call) + ": " + nccl_api().GetErrorString(r__);        \
            return PK_E_CUDA;                                                                    \
        }                                                                                        \
    } while (0)

int pk_comm_get_id(pk_comm_id *id)
{
    if (!id) return PK_E_INVALID;
    static_assert(sizeof(pk_comm_id) == sizeof(ncclUniqueId), "pk_comm_id must hold an ncclUniqueId");
    NcclApi &api = nccl_api();
    if (!api.error.empty()) return PK_E_STATE;
    ncclUniqueId u;
    if (api.GetUniqueId(&u) != ncclSuccess) return PK_E_CUDA;
    std::memcpy(id, &u, sizeof(u));
    return PK_OK;
}

int pk_comm_init(pk_ctx *ctx, const pk_comm_id *id, int rank, int nranks)
{
    if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return PK_E_INVALID;
    if (ctx->comm) return PK_E_STATE;
    NcclApi &api = nccl_api();
    if (!api.error.empty())
    {
        ctx->last_error = api.error;
        return PK_E_STATE;
    }
    cudaSetDevice(ctx->cfg.device);
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    PK_NCCL(api.CommInitRank(&ctx->comm, nranks, u, rank));
    ctx->comm_rank = rank;
    ctx->comm_size = nranks;
    PK_CUDA(cudaMalloc(reinterpret_cast<void **>(&ctx->d_comm_counts), (static_cast<size_t>(nranks) + 1) * sizeof(unsigned long long)));
    PK_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_comm_counts), (static_cast<size_t>(nranks) + 1) * sizeof(unsigned long long),
                          cudaHostAllocDefault));
    for (auto &e : ctx->ev_comm) PK_CUDA(cudaEventCreate(&e));
    return PK_OK;
}

int pk_comm_pose_slice(pk_ctx *ctx, uint32_t *first, uint32_t *count)
{
    if (!ctx || !first || !count) return PK_E_INVALID;
    const uint64_t n = ctx->n_bodies, N = static_cast<uint64_t>(ctx->comm_size), r = static_cast<uint64_t>(ctx->comm_rank);
    *first = static_cast<uint32_t>(n * r / N);
    *count = static_cast<uint32_t>(n * (r + 1) / N - n * r / N);
    return PK_OK;
}

int pk_comm_allgather_poses(pk_ctx *ctx, int what)
{
    if (!ctx) return PK_E_INVALID;
    PkRange range("pk_comm_allgather_poses");
    if (!ctx->comm) return PK_E_STATE;
    cudaSetDevice(ctx->cfg.device);
    NcclApi &api = nccl_api();
    const uint64_t n = ctx->n_bodies, N = static_cast<uint64_t>(ctx->comm_size);
    struct Arr
    {
        double *p;
        uint64_t width;
        int bit;
    } arrs[3] = {{ctx->d_pos, 3, PK_POSE_POS}, {ctx->d_quat, 4, PK_POSE_QUAT}, {ctx->d_disp, 3, PK_POSE_DISP}};
    for (const Arr &a : arrs)
    {
        if (!(what & a.bit)) continue;
        if (n % N == 0)
        {
            const uint64_t chunk = n / N * a.width; // in place: the send block is this rank's block of the receive buffer
            PK_NCCL(api.AllGather(a.p + chunk * static_cast<uint64_t>(ctx->comm_rank), a.p, chunk, ncclDouble, ctx->comm, ctx->stream));
        }
        else
        {
            PK_NCCL(api.GroupStart());
            for (uint64_t r = 0; r < N; ++r)
            {
                const uint64_t f = n * r / N, c = n * (r + 1) / N - f;
                if (c) PK_NCCL(api.Broadcast(a.p + f * a.width, a.p + f * a.width, c * a.width, ncclDouble, static_cast<int>(r), ctx->comm, ctx->stream));
            }
            PK_NCCL(api.GroupEnd());
        }
    }
    if ((what & PK_POSE_QUAT) && ctx->dyn_enabled)
        dynamics_derive_kernel<<<div_up(ctx->n_bodies, 256), 256, 0, ctx->stream>>>(ctx->d_quat, ctx->dyn, 0, ctx->n_bodies);
    return PK_OK;
}

int pk_comm_allgather_contacts(pk_ctx *ctx, pk_gathered_contacts *out)
{
    if (!ctx || !out) return PK_E_INVALID;
    PkRange range("pk_comm_allgather_contacts");
    if (!ctx->comm) return PK_E_STATE;
    if (!ctx->have_results || !ctx->device_results) return PK_E_STATE;
    cudaSetDevice(ctx->cfg.device);
    NcclApi &api = nccl_api();
    cudaStream_t s = ctx->stream;
    const int N = ctx->comm_size;
    cudaEventRecord(ctx->ev_comm[0], s);
    // Every rank's count (a one-word all-gather straight from the step's counter) and the records, straight from the
    // contact buffer the EPA kernels wrote (no staging copy), in blocks of `stride` records; what a block holds beyond
    // its rank's count is unspecified.  The block size must be the same on every rank and known before the counts are:
    // it is taken from the previous exchange — the largest count seen then, plus an eighth — and both all-gathers go
    // out as one group with one synchronisation at the end.  Only when some rank outgrew the guess (or on the first
    // exchange) are the records gathered a second time, in blocks of the largest count.
    auto ensure_gather = [&](uint64_t records) -> int
    {
        if (records <= ctx->gather_stride) return PK_OK;
        if (ctx->d_gather) cudaFree(ctx->d_gather);
        ctx->d_gather = nullptr;
        ctx->gather_stride = 0;
        const uint64_t cap = std::min<uint64_t>(ctx->max_contacts, records + records / 4 + 1024);
        PK_CUDA(cudaMalloc(reinterpret_cast<void **>(&ctx->d_gather), static_cast<size_t>(N) * cap * sizeof(ContactRec)));
        ctx->gather_stride = cap;
        return PK_OK;
    };
    const uint64_t guess = ctx->gather_guess ? std::min<uint64_t>(ctx->max_contacts, ctx->gather_guess + ctx->gather_guess / 8 + 1024) : 0;
    if (guess) PK_TRY(ensure_gather(guess));
    ctx->h_comm_counts[N] = ctx->num_contacts;
    PK_CUDA(cudaMemcpyAsync(ctx->d_comm_counts + N, ctx->h_comm_counts + N, sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
    PK_NCCL(api.GroupStart());
    PK_NCCL(api.AllGather(ctx->d_comm_counts + N, ctx->d_comm_counts, 1, ncclUint64, ctx->comm, s));
    if (guess) PK_NCCL(api.AllGather(ctx->d_contacts_final, ctx->d_gather, guess * sizeof(ContactRec), ncclChar, ctx->comm, s));
    PK_NCCL(api.GroupEnd());
    PK_CUDA(cudaMemcpyAsync(ctx->h_comm_counts, ctx->d_comm_counts, static_cast<size_t>(N) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    cudaEventRecord(ctx->ev_comm[1], s);
    PK_CUDA(cudaStreamSynchronize(s));
    uint64_t stride = 0, total = 0;
    for (int r = 0; r < N; ++r)
    {
        stride = std::max<uint64_t>(stride, ctx->h_comm_counts[r]);
        total += ctx->h_comm_counts[r];
    }
    if (stride > ctx->max_contacts)
    {
        ctx->last_error = "a rank holds more contacts than this context's max_contacts: ranks must be created with equal capacities";
        return PK_E_STATE;
    }
    ctx->gather_guess = stride;
    if (stride > guess)
    {
        PK_TRY(ensure_gather(stride));
        PK_NCCL(api.AllGather(ctx->d_contacts_final, ctx->d_gather, stride * sizeof(ContactRec), ncclChar, ctx->comm, s));
        cudaEventRecord(ctx->ev_comm[1], s);
        PK_CUDA(cudaStreamSynchronize(s));
    }
    else
        stride = guess; // the blocks that were gathered
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev_comm[0], ctx->ev_comm[1]);
    out->d_records = ctx->d_gather;
    out->stride_records = stride;
    out->counts = reinterpret_cast<const uint64_t *>(ctx->h_comm_counts);
    out->num_ranks = static_cast<uint32_t>(N);
    out->total = total;
    out->ms = ms;
    return PK_OK;
}

// ------------------------------------------------------------------------------------ gjk_epa batch

int pk_gjk_epa_batch_device(pk_ctx *ctx, const uint32_t *d_a, const uint32_t *d_b, uint64_t n, pk_contact *d_out,
                            uint8_t *d_hit, float *ms)
{
    if (!ctx || !d_a || !d_b || !d_out || !d_hit) return PK_E_INVALID;
    if (n > ctx->cfg.max_pairs) return PK_E_PAIR_OVERFLOW;
    cudaSetDevice(ctx->cfg.device);
    PK_TRY(upload_shapes(ctx));
    cudaStream_t s = ctx->stream;
    ctx->launches = 0;
    cudaEventRecord(ctx->ev[ST_BOUNDS], s);
    // the batch shares the narrowphase buffers (hit flags, contact slots, contact records) with the step: what the
    // last pk_collide left there is gone: results that were not fetched to the host can no longer be, nor merged
    // into manifolds (PK_E_STATE).  Host copies already fetched, the pair keys, the tree and the scene box of that
    // step (pk_raycast) stay valid.
    ctx->device_results = false;
    scene_reset_kernel<<<1, 64, 0, s>>>(nullptr, ctx->d_counters, C_COUNT);
    PK_TRY(run_narrowphase(ctx, nullptr, d_a, d_b, n, true));
    if (n)
        expand_contacts_kernel<<<div_up(n, 256), 256, 0, s>>>(ctx->d_hit, ctx->d_out_index, ctx->d_valid, ctx->d_contacts[0],
                                                             d_a, d_b, n, ctx->max_contacts, reinterpret_cast<ContactRec *>(d_out), d_hit);
    ctx->launches += 2;
    cudaEventRecord(ctx->ev[ST_COUNT], s);
    PK_TRY(read_counters(ctx));
    PK_CUDA(cudaGetLastError());
    if (ms) cudaEventElapsedTime(ms, ctx->ev[ST_BOUNDS], ctx->ev[ST_COUNT]);
    cudaEventElapsedTime(&ctx->stage_ms[ST_GJK], ctx->ev[ST_GJK], ctx->ev[ST_SCAN]);
    cudaEventElapsedTime(&ctx->stage_ms[ST_SCAN], ctx->ev[ST_SCAN], ctx->ev[ST_EPA]);
    cudaEventElapsedTime(&ctx->stage_ms[ST_EPA], ctx->ev[ST_EPA], ctx->ev[ST_COMPACT]);
    if (ctx->h_counters[C_HITS] > ctx->max_contacts)
    {
        ctx->last_error = "GJK hits exceed pk_config.max_contacts";
        return PK_E_PAIR_OVERFLOW;
    }
    if (ctx->h_counters[C_EPA_OVERFLOW])
    {
        ctx->last_error = "EPA polytope exceeded the per-pair scratch for some pairs";
        return PK_E_EPA_OVERFLOW;
    }
    return PK_OK;
}

int pk_gjk_epa_batch(pk_ctx *ctx, const uint32_t *pair_a, const uint32_t *pair_b, uint64_t n, pk_contact *out, uint8_t *hit)
{
    if (!ctx || !pair_a || !pair_b || !out || !hit) return PK_E_INVALID;
    if (n > ctx->cfg.max_pairs) return PK_E_PAIR_OVERFLOW;
    if (n == 0) return PK_OK;
    for (uint64_t k = 0; k < n; ++k)
        if (pair_a[k] >= ctx->n_bodies || pair_b[k] >= ctx->n_bodies) return PK_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    uint32_t *d_a = nullptr, *d_b = nullptr;
    ContactRec *d_out = nullptr;
    uint8_t *d_h = nullptr;
    int status = PK_OK;
    auto cleanup = [&]()
    {
        if (d_a) cudaFree(d_a);
        if (d_b) cudaFree(d_b);
        if (d_out) cudaFree(d_out);
        if (d_h) cudaFree(d_h);
    };
    if ((status = dev_alloc(ctx, &d_a, n)) != PK_OK || (status = dev_alloc(ctx, &d_b, n)) != PK_OK ||
        (status = dev_alloc(ctx, &d_out, n)) != PK_OK || (status = dev_alloc(ctx, &d_h, n)) != PK_OK)
    {
        cleanup();
        return status;
    }
    cudaMemcpyAsync(d_a, pair_a, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_b, pair_b, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    status = pk_gjk_epa_batch_device(ctx, d_a, d_b, n, reinterpret_cast<pk_contact *>(d_out), d_h, nullptr);
    if (status == PK_OK || status == PK_E_EPA_OVERFLOW)
    {
        cudaMemcpyAsync(out, d_out, n * sizeof(pk_contact), cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(hit, d_h, n, cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) status = PK_E_CUDA;
    }
    cleanup();
    return status;
}

// ------------------------------------------------------------------------------------ gjk distance batch
static_assert(sizeof(pk_distance) == sizeof(DistanceRec), "pk_distance layout");

int pk_gjk_distance_batch_device(pk_ctx *ctx, const uint32_t *d_a, const uint32_t *d_b, uint64_t n, pk_distance *d_out,
                                 uint8_t *d_separated, float *ms)
{
    if (!ctx || !d_a || !d_b || !d_out || !d_separated) return PK_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    PK_TRY(upload_shapes(ctx));
    cudaStream_t s = ctx->stream;
    // own events: the stage events of the step (pk_get_report) are left alone, like every buffer of the step
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ms)
    {
        PK_CUDA(cudaEventCreate(&e0));
        if (cudaEventCreate(&e1) != cudaSuccess)
        {
            cudaEventDestroy(e0);
            return PK_E_CUDA;
        }
        cudaEventRecord(e0, s);
    }
    if (n)
    {
        nvtxRangePushA("pk:gjk_distance");
        auto kern = ctx->has_big_hulls ? gjk_distance_kernel<true> : gjk_distance_kernel<false>;
        kern<<<static_cast<unsigned>(div_up(n, 128)), 128, 0, s>>>(body_arrays(ctx), d_a, d_b, n, ctx->n_bodies,
                                                                     reinterpret_cast<DistanceRec *>(d_out), d_separated);
        nvtxRangePop();
    }
    int status = PK_OK;
    if (ms)
    {
        cudaEventRecord(e1, s);
        if (cudaEventSynchronize(e1) != cudaSuccess || cudaEventElapsedTime(ms, e0, e1) != cudaSuccess) status = PK_E_CUDA;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    else if (cudaStreamSynchronize(s) != cudaSuccess)
        status = PK_E_CUDA;
    if (cudaGetLastError() != cudaSuccess) status = PK_E_CUDA;
    return status;
}

int pk_gjk_distance_batch(pk_ctx *ctx, const uint32_t *pair_a, const uint32_t *pair_b, uint64_t n, pk_distance *out, uint8_t *separated)
{
    if (!ctx) return PK_E_INVALID;
    if (n == 0) return PK_OK;
    if (!pair_a || !pair_b || !out || !separated) return PK_E_INVALID;
    for (uint64_t k = 0; k < n; ++k)
        if (pair_a[k] >= ctx->n_bodies || pair_b[k] >= ctx->n_bodies) return PK_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    uint32_t *d_a = nullptr, *d_b = nullptr;
    DistanceRec *d_out = nullptr;
    uint8_t *d_s = nullptr;
    int status = PK_OK;
    auto cleanup = [&]()
    {
        if (d_a) cudaFree(d_a);
        if (d_b) cudaFree(d_b);
        if (d_out) cudaFree(d_out);
        if (d_s) cudaFree(d_s);
    };
    if ((status = dev_alloc(ctx, &d_a, n)) != PK_OK || (status = dev_alloc(ctx, &d_b, n)) != PK_OK ||
        (status = dev_alloc(ctx, &d_out, n)) != PK_OK || (status = dev_alloc(ctx, &d_s, n)) != PK_OK)
    {
        cleanup();
        return status;
    }
    cudaMemcpyAsync(d_a, pair_a, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_b, pair_b, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    status = pk_gjk_distance_batch_device(ctx, d_a, d_b, n, reinterpret_cast<pk_distance *>(d_out), d_s, nullptr);
    if (status == PK_OK)
    {
        cudaMemcpyAsync(out, d_out, n * sizeof(pk_distance), cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(separated, d_s, n, cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) status = PK_E_CUDA;
    }
    cleanup();
    return status;
}

// ------------------------------------------------------------------------------------ raw memory helpers
int pk_device_alloc(pk_ctx *ctx, size_t bytes, void **dptr)
{
    if (!ctx || !dptr) return PK_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    PK_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
    return PK_OK;
}
int pk_device_free(pk_ctx *ctx, void *dptr)
{
    if (!ctx) return PK_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    PK_CUDA(cudaFree(dptr));
    return PK_OK;
}
int pk_memcpy_h2d(pk_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx) return PK_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    PK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    return PK_OK;
}
int pk_memcpy_d2h(pk_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx) return PK_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    PK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    return PK_OK;
}

int pk_memcpy_d2d(pk_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx) return PK_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    PK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    return PK_OK;
}
int pk_host_alloc(pk_ctx *ctx, size_t bytes, void **hptr)
{
    if (!ctx || !hptr) return PK_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    PK_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return PK_OK;
}
int pk_host_free(pk_ctx *ctx, void *hptr)
{
    if (!ctx) return PK_E_INVALID;
    PK_CUDA(cudaFreeHost(hptr));
    return PK_OK;
}

} // extern "C"
