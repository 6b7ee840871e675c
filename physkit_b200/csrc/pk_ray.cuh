// pk_ray.cuh — batched world ray casts over the step's LBVH.
//
// Reproduces world_base::raycast (core/world.h:260-319) = dynamic_bvh::raycast (bvh.h:346-450) over the
// static and the dynamic tree: every body whose STORED (fat) box the ray enters within max_distance, with
// the entry distance of ray::intersect_distance (bvh.h:59-98).  The reference yields them in an order that
// depends on its trees' shape; the contract here is the set, delivered sorted by (ray, body), and — in
// closest mode — the entry with the smallest distance (lowest body id among equals), which is what the
// callback form computes when the callback shrinks max_distance (bvh.h:367-373).
//
// One thread per ray, stackless rope traversal of the tree pk_collide built.  Internal nodes hold float
// boxes rounded outward; they are tested with the reference's FP64 slab expression on the float box
// widened to double.  That test is conservative: every step of the expression is monotone in the box
// ((lo − o)·inv, fmin/fmax, max/min, the NaN → ±inf substitution), so a box that contains a leaf's stored
// box passes whenever the leaf's own test passes.  The decision and the distance come from the exact
// 64-byte leaf record.  Worlds of a batched context are tiled in float space (pk_broadphase.cuh): the node
// box is moved back by the world's offset, one ulp outward, and the leaf test requires the ray's world.
#pragma once

#include "pk_broadphase.cuh"

namespace pk
{

constexpr int RAY_THREADS = 128;
constexpr int RAY_MODE_ALL = 0;
constexpr int RAY_MODE_CLOSEST = 1;

struct RayQ
{
    d3 o, inv;
};

// ray::intersect_distance (bvh.h:59-98), operation for operation.
__device__ __forceinline__ bool ray_box(const RayQ &r, double lox, double loy, double loz, double hix, double hiy, double hiz,
                                        double max_distance, double &dist)
{
    const double inf = __longlong_as_double(0x7FF0000000000000ll);
    double ax = (lox - r.o.x) * r.inv.x, bx = (hix - r.o.x) * r.inv.x;
    double ay = (loy - r.o.y) * r.inv.y, by = (hiy - r.o.y) * r.inv.y;
    double az = (loz - r.o.z) * r.inv.z, bz = (hiz - r.o.z) * r.inv.z;
    if (ax != ax) ax = -inf;
    if (bx != bx) bx = inf;
    if (ay != ay) ay = -inf;
    if (by != by) by = inf;
    if (az != az) az = -inf;
    if (bz != bz) bz = inf;
    const double tminx = fmin(ax, bx), tmaxx = fmax(ax, bx);
    const double tminy = fmin(ay, by), tmaxy = fmax(ay, by);
    const double tminz = fmin(az, bz), tmaxz = fmax(az, bz);
    // std::max({a,b,c}) / std::min({a,b,c}): no NaN can reach them, so the value is the plain maximum
    const double tmin = dmax(dmax(tminx, tminy), tminz);
    const double tmax = dmin(dmin(tmaxx, tmaxy), tmaxz);
    if (tmax >= 0.0 && tmin <= tmax && tmin <= max_distance)
    {
        dist = dmax(0.0, tmin); // origin inside the box: clamp to 0
        return true;
    }
    return false;
}

__device__ __forceinline__ double next_down(double x) // x − 1 ulp for finite x (toward −inf)
{
    if (x == 0.0) return -4.9406564584124654e-324;
    long long b = __double_as_longlong(x);
    return __longlong_as_double(x > 0.0 ? b - 1 : b + 1);
}
__device__ __forceinline__ double next_up(double x) { return -next_down(-x); }

__global__ void __launch_bounds__(RAY_THREADS)
ray_cast_kernel(const NodeF *__restrict__ nodes, const LeafRec *__restrict__ leaves, uint32_t m, const uint32_t *__restrict__ root_ptr,
                const double *__restrict__ origins, const double *__restrict__ dirs, const double *__restrict__ max_dist,
                const uint32_t *__restrict__ ray_world, uint32_t nrays, int mode, WorldTiling wt, const uint32_t *__restrict__ scene,
                uint64_t *__restrict__ out_keys, double *__restrict__ out_dist, uint32_t *__restrict__ out_src, uint64_t capacity,
                unsigned long long *__restrict__ counter)
{
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const uint32_t ray = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool live = ray < nrays;
    RayQ q{};
    double limit = 0.0;
    uint32_t world = 0;
    double off[3] = {0.0, 0.0, 0.0};
    uint32_t node = NODE_SENTINEL;
    const uint32_t first_leaf = m - 1;
    if (live)
    {
        q.o = d3{origins[3ull * ray], origins[3ull * ray + 1], origins[3ull * ray + 2]};
        const d3 dir = normalized(d3{dirs[3ull * ray], dirs[3ull * ray + 1], dirs[3ull * ray + 2]}); // ray ctor, bvh.h:47-50
        q.inv = d3{1.0 / dir.x, 1.0 / dir.y, 1.0 / dir.z};                                             // safe_inv, bvh.h:27-35
        limit = max_dist[ray];
        world = ray_world ? ray_world[ray] : 0u;
        if (wt.num_worlds > 1)
        {
            float fo[3];
            world_offset(wt, world, scene_tile(scene), fo);
            off[0] = fo[0];
            off[1] = fo[1];
            off[2] = fo[2];
        }
        node = (m >= 2) ? *root_ptr : first_leaf; // a single leaf is its own tree
    }
    const bool tiled = wt.num_worlds > 1;
    uint32_t best_id = 0xFFFFFFFFu;
    double best_d = 0.0;
    while (__any_sync(FULL, node != NODE_SENTINEL))
    {
        bool emit = false;
        uint32_t hit_id = 0;
        double hit_d = 0.0;
        if (node != NODE_SENTINEL)
        {
            const float4 *np = reinterpret_cast<const float4 *>(nodes + node);
            const float4 n0 = __ldg(np), n1 = __ldg(np + 1);
            double lox = n0.x, loy = n0.y, loz = n0.z, hix = n0.w, hiy = n1.x, hiz = n1.y;
            if (tiled)
            {
                lox = next_down(lox - off[0]);
                loy = next_down(loy - off[1]);
                loz = next_down(loz - off[2]);
                hix = next_up(hix - off[0]);
                hiy = next_up(hiy - off[1]);
                hiz = next_up(hiz - off[2]);
            }
            double dn;
            uint32_t next = __float_as_uint(n1.w);
            if (ray_box(q, lox, loy, loz, hix, hiy, hiz, limit, dn))
            {
                if (node >= first_leaf)
                {
                    const double2 *lp = reinterpret_cast<const double2 *>(leaves + (node - first_leaf));
                    const double2 b0 = __ldg(lp), b1 = __ldg(lp + 1), b2 = __ldg(lp + 2);
                    const int4 meta = __ldg(reinterpret_cast<const int4 *>(lp + 3));
                    double d;
                    if (static_cast<uint32_t>(meta.w) == world && ray_box(q, b0.x, b0.y, b1.x, b1.y, b2.x, b2.y, limit, d))
                    {
                        const uint32_t id = static_cast<uint32_t>(meta.x);
                        if (mode == RAY_MODE_ALL)
                        {
                            emit = true;
                            hit_id = id;
                            hit_d = d;
                        }
                        else if (best_id == 0xFFFFFFFFu || d < best_d || (d == best_d && id < best_id))
                        {
                            best_id = id;
                            best_d = d;
                            limit = d; // the callback's new max_distance (bvh.h:371): inclusive, ties are still visited
                        }
                    }
                }
                else
                    next = __float_as_uint(n1.z); // left child
            }
            node = next;
        }
        if (mode == RAY_MODE_ALL)
        {
            const unsigned em = __ballot_sync(FULL, emit);
            if (em)
            {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(counter, static_cast<unsigned long long>(__popc(em)));
                base = __shfl_sync(FULL, base, 0);
                const unsigned long long at = base + __popc(em & ((1u << lane) - 1u));
                if (emit && at < capacity)
                {
                    out_keys[at] = (static_cast<uint64_t>(ray) << 32) | hit_id;
                    out_dist[at] = hit_d;
                    out_src[at] = static_cast<uint32_t>(at);
                }
            }
        }
    }
    if (mode == RAY_MODE_CLOSEST)
    {
        const bool emit = live && best_id != 0xFFFFFFFFu;
        const unsigned em = __ballot_sync(FULL, emit);
        if (em)
        {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(counter, static_cast<unsigned long long>(__popc(em)));
            base = __shfl_sync(FULL, base, 0);
            const unsigned long long at = base + __popc(em & ((1u << lane) - 1u));
            if (emit && at < capacity)
            {
                out_keys[at] = (static_cast<uint64_t>(ray) << 32) | best_id;
                out_dist[at] = best_d;
                out_src[at] = static_cast<uint32_t>(at);
            }
        }
    }
}

// Fewer than two bodies alive: the step builds no tree; the ray is tested against the stored boxes directly.
__global__ void __launch_bounds__(RAY_THREADS)
ray_brute_kernel(const double *__restrict__ stored, const uint8_t *__restrict__ alive, const uint32_t *__restrict__ body_world,
                 uint32_t nbodies, const double *__restrict__ origins, const double *__restrict__ dirs,
                 const double *__restrict__ max_dist, const uint32_t *__restrict__ ray_world, uint32_t nrays, int mode,
                 uint64_t *__restrict__ out_keys, double *__restrict__ out_dist, uint32_t *__restrict__ out_src, uint64_t capacity,
                 unsigned long long *__restrict__ counter)
{
    const uint32_t ray = blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= nrays) return;
    RayQ q;
    q.o = d3{origins[3ull * ray], origins[3ull * ray + 1], origins[3ull * ray + 2]};
    const d3 dir = normalized(d3{dirs[3ull * ray], dirs[3ull * ray + 1], dirs[3ull * ray + 2]});
    q.inv = d3{1.0 / dir.x, 1.0 / dir.y, 1.0 / dir.z};
    double limit = max_dist[ray];
    const uint32_t world = ray_world ? ray_world[ray] : 0u;
    uint32_t best_id = 0xFFFFFFFFu;
    double best_d = 0.0;
    for (uint32_t i = 0; i < nbodies; ++i)
    {
        if (!alive[i] || (body_world && body_world[i] != world)) continue;
        const double *b = stored + 6ull * i;
        double d;
        if (!ray_box(q, b[0], b[1], b[2], b[3], b[4], b[5], limit, d)) continue;
        if (mode == RAY_MODE_ALL)
        {
            const unsigned long long at = atomicAdd(counter, 1ull);
            if (at < capacity)
            {
                out_keys[at] = (static_cast<uint64_t>(ray) << 32) | i;
                out_dist[at] = d;
                out_src[at] = static_cast<uint32_t>(at);
            }
        }
        else if (best_id == 0xFFFFFFFFu || d < best_d)
        {
            best_id = i;
            best_d = d;
            limit = d;
        }
    }
    if (mode == RAY_MODE_CLOSEST && best_id != 0xFFFFFFFFu)
    {
        const unsigned long long at = atomicAdd(counter, 1ull);
        if (at < capacity)
        {
            out_keys[at] = (static_cast<uint64_t>(ray) << 32) | best_id;
            out_dist[at] = best_d;
            out_src[at] = static_cast<uint32_t>(at);
        }
    }
}

// 16-byte result record, mirrors pk_ray_hit of include/pk_collide.h
struct RayHitRec
{
    uint32_t ray;
    uint32_t body;
    double distance;
};
static_assert(sizeof(RayHitRec) == 16, "RayHitRec must match pk_ray_hit");

__global__ void __launch_bounds__(256)
ray_gather_kernel(const uint64_t *__restrict__ sorted_keys, const uint32_t *__restrict__ sorted_src, const double *__restrict__ dist,
                  uint64_t n, RayHitRec *__restrict__ out)
{
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = sorted_keys[i];
    RayHitRec r;
    r.ray = static_cast<uint32_t>(k >> 32);
    r.body = static_cast<uint32_t>(k & 0xFFFFFFFFu);
    r.distance = dist[sorted_src[i]];
    out[i] = r;
}

} // namespace pk
