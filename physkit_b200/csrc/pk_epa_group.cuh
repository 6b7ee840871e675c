// pk_epa_group.cuh — EPA (reference src/collision.cpp:251-509) with one lane GROUP per pair and the
// polytope resident in shared memory.
//
// Why: the thread-per-pair kernel (pk_narrowphase.cuh, epa_kernel) keeps each polytope in a 52 KB HBM
// slab; 56 832 resident polytopes do not fit the L2, so every dependent access of the face heap / flood
// fill is a DRAM round trip (ncu r1: 71 % long_scoreboard stalls, 31 GB DRAM traffic for 0.4 GB of
// algorithmic bytes, 100 k cycles per EPA iteration per thread).  Here a group of G lanes (4 or 8) owns
// one pair:
//   * plane / topology of the first F_S live faces, the first H_S heap entries and the first V_S vertex
//     positions live in the group's private shared-memory block (≈4.7 KB → 48 pairs per SM); anything
//     beyond spills to a small per-group HBM area (rare: only polytopes past ≈30 iterations)
//   * face slots are recycled (a face made obsolete frees its slot; heap entries carry the face's creation
//     serial so lazily deleted entries are still recognised), which keeps the live set at 2V−4 faces
//   * sequential parts of the algorithm (heap sifts, the LIFO flood fill's bookkeeping) run uniformly on
//     all lanes of the group; the parallel parts use the lanes: the two support functions (even lanes
//     shape A, odd lanes shape B, hull scans split over the lanes of a parity), the three neighbour
//     visibility tests of a flood-fill step, one new face per lane, ring links through a vertex→edge table
//   * groups never synchronise with each other: all exchanges are __shfl/__syncwarp on the group's mask
//
// Results are bit-identical to the thread-per-pair kernel (same arithmetic, same heap order, same
// horizon order).  Slot recycling relies on the polytope being a proper manifold (every link matched,
// horizon a set of simple loops); the moment a pair violates that, needs a padded simplex, or outgrows
// the hard capacities, it is appended to a fallback list that epa_kernel processes afterwards with the
// reference's own bookkeeping.
#pragma once

#include "pk_narrowphase.cuh"

namespace pk
{

constexpr int EG_MAX_SLOTS = 136; // live faces ≤ 2·68 − 4 = 132
constexpr int EG_MAX_HEAP = EPA_MAX_FACES;
constexpr int EG_MAX_HORIZON = 32;

template <int G, int F_S, int H_S, int V_S> struct alignas(16) EgSmem
{
    double plane[F_S][4];         // normal xyz, distance
    unsigned long long topo[F_S]; // bytes 0-2 vertices, 3-5 adjacent slots (0xFF = none), 6-7 creation serial (0xFFFF = free)
    double hdist[H_S + 2];        // heap, node j in [1, size] (children 2j, 2j+1 share one 16-byte load): face distance
    uint32_t hid[H_S + 2];        // heap: serial << 16 | slot
    double vpos[V_S][3];          // p = pa − pb per polytope vertex
    uint8_t hz_start[EG_MAX_HORIZON], hz_end[EG_MAX_HORIZON], hz_adj[EG_MAX_HORIZON], hz_slot[EG_MAX_HORIZON];
    // groups of one warp touch the same field at the same time: the stride is chosen so that they land in
    // different banks (16 bytes apart for 8 groups per warp, 32 for 4)
    static constexpr int RAW = F_S * 40 + (H_S + 2) * 12 + V_S * 24 + 4 * EG_MAX_HORIZON + 144;
    static constexpr int WANT = (G == 4) ? 16 : 32;
    static constexpr int PAD = ((WANT - RAW % 128) + 128) % 128;
    uint8_t edge_tables[144 + PAD]; // vertex → horizon edge starting / ending there (0xFF = none), then padding
    __device__ __forceinline__ uint8_t *edge_of_start() { return edge_tables; }
    __device__ __forceinline__ uint8_t *edge_of_end() { return edge_tables + 72; }
};

template <int F_S, int H_S, int V_S> struct EgSpillLayout
{
    static constexpr size_t PLANE = 0;
    static constexpr size_t TOPO = PLANE + static_cast<size_t>(EG_MAX_SLOTS - F_S) * 32;
    static constexpr size_t VPOS = (TOPO + static_cast<size_t>(EG_MAX_SLOTS - F_S) * 8 + 15) / 16 * 16;
    static constexpr size_t VAB = VPOS + static_cast<size_t>(EPA_MAX_VERTS - V_S) * 24;
    static constexpr size_t BYTES = (VAB + static_cast<size_t>(EPA_MAX_VERTS) * 48 + 127) / 128 * 128;
};

// Group view of the polytope: shared-memory part + spill part.
template <int G, int F_S, int H_S, int V_S> struct EgPoly
{
    EgSmem<G, F_S, H_S, V_S> *s;
    double *g_plane;
    unsigned long long *g_topo;
    double *g_vpos;
    double *g_vab; // pa xyz, pb xyz per vertex (cold: only the result needs it)

    __device__ __forceinline__ double4 plane(int f) const
    {
        double2 a, b;
        if (f < F_S)
        {
            const double2 *q = reinterpret_cast<const double2 *>(s->plane[f]);
            a = q[0];
            b = q[1];
        }
        else
        {
            const double2 *q = reinterpret_cast<const double2 *>(g_plane + 4 * (f - F_S));
            a = q[0];
            b = q[1];
        }
        return make_double4(a.x, a.y, b.x, b.y);
    }
    __device__ __forceinline__ double plane_dist(int f) const { return (f < F_S) ? s->plane[f][3] : g_plane[4 * (f - F_S) + 3]; }
    __device__ __forceinline__ void set_plane(int f, d3 n, double dist) const
    {
        if (f < F_S)
        {
            double2 *q = reinterpret_cast<double2 *>(s->plane[f]);
            q[0] = make_double2(n.x, n.y);
            q[1] = make_double2(n.z, dist);
        }
        else
        {
            double2 *q = reinterpret_cast<double2 *>(g_plane + 4 * (f - F_S));
            q[0] = make_double2(n.x, n.y);
            q[1] = make_double2(n.z, dist);
        }
    }
    __device__ __forceinline__ unsigned long long topo(int f) const { return (f < F_S) ? s->topo[f] : g_topo[f - F_S]; }
    __device__ __forceinline__ void set_topo(int f, unsigned long long t) const
    {
        if (f < F_S)
            s->topo[f] = t;
        else
            g_topo[f - F_S] = t;
    }
    __device__ __forceinline__ void set_adj(int f, int e, int to) const
    {
        if (f < F_S)
            reinterpret_cast<uint8_t *>(&s->topo[f])[3 + e] = static_cast<uint8_t>(to);
        else
            reinterpret_cast<uint8_t *>(&g_topo[f - F_S])[3 + e] = static_cast<uint8_t>(to);
    }
    __device__ __forceinline__ d3 vpos(int i) const
    {
        if (i < V_S) return {s->vpos[i][0], s->vpos[i][1], s->vpos[i][2]};
        const double *q = g_vpos + 3 * (i - V_S);
        return {q[0], q[1], q[2]};
    }
    __device__ __forceinline__ void set_vpos(int i, d3 p) const
    {
        if (i < V_S)
        {
            s->vpos[i][0] = p.x;
            s->vpos[i][1] = p.y;
            s->vpos[i][2] = p.z;
        }
        else
        {
            double *q = g_vpos + 3 * (i - V_S);
            q[0] = p.x;
            q[1] = p.y;
            q[2] = p.z;
        }
    }
};

__device__ __forceinline__ int eg_v(unsigned long long t, int i) { return static_cast<int>((t >> (8 * i)) & 0xFFull); }
__device__ __forceinline__ int eg_adj(unsigned long long t, int i) { return static_cast<int>((t >> (24 + 8 * i)) & 0xFFull); }
__device__ __forceinline__ uint32_t eg_serial(unsigned long long t) { return static_cast<uint32_t>(t >> 48); }
constexpr unsigned long long EG_DEAD = 0xFFFFull << 48;

// The face heap (collision.cpp:390-408) restated from libstdc++ with 1-based node numbers j = k + 1
// (parent j/2, children 2j and 2j+1), entirely in shared memory.  Every lane of the group executes this
// with identical operands (identical stores to one address merge).
// std::__push_heap with comp(a,b) = dist[a] > dist[b]: move the hole up while the parent is farther.
template <class SM> __device__ __forceinline__ void eg_sift_up(SM *s, int j, double vd, uint32_t vi)
{
    while (j > 1)
    {
        const int pj = j >> 1;
        const double pd = s->hdist[pj];
        if (!(pd > vd)) break;
        s->hdist[j] = pd;
        s->hid[j] = s->hid[pj];
        j = pj;
    }
    s->hdist[j] = vd;
    s->hid[j] = vi;
}
template <class SM> __device__ __forceinline__ void eg_heap_push(SM *s, int &size, double vd, uint32_t vi)
{
    ++size;
    eg_sift_up(s, size, vd, vi);
}
// std::pop_heap (→ __pop_heap → __adjust_heap: hole to the bottom along the closer child, right child on
// ties, then __push_heap of the former last element) followed by back()/pop_back()
template <class SM> __device__ __forceinline__ uint32_t eg_heap_pop(SM *s, int &size)
{
    const uint32_t top = s->hid[1];
    if (size == 1)
    {
        size = 0;
        return top;
    }
    const int len = size - 1;
    const double vd = s->hdist[size];
    const uint32_t vi = s->hid[size];
    int j = 1;
    while (j - 1 < (len - 1) / 2)
    {
        const int l = 2 * j;
        const double2 cd = *reinterpret_cast<const double2 *>(&s->hdist[l]);
        const uint2 ci = *reinterpret_cast<const uint2 *>(&s->hid[l]);
        const bool left = cd.y > cd.x;
        s->hdist[j] = left ? cd.x : cd.y;
        s->hid[j] = left ? ci.x : ci.y;
        j = left ? l : l + 1;
    }
    if ((len & 1) == 0 && j - 1 == (len - 2) / 2)
    {
        const int l = 2 * j;
        s->hdist[j] = s->hdist[l];
        s->hid[j] = s->hid[l];
        j = l;
    }
    eg_sift_up(s, j, vd, vi);
    size = len;
    return top;
}

__device__ __forceinline__ d3 eg_shfl_xor(unsigned mask, d3 v, int x)
{
    return {__shfl_xor_sync(mask, v.x, x), __shfl_xor_sync(mask, v.y, x), __shfl_xor_sync(mask, v.z, x)};
}

// Support of the lane's shape along d.  Analytic shapes: every lane of the parity computes the same
// value.  Hulls (src/mesh.cpp:341-358): the NSUB lanes of the parity scan interleaved vertex subsets and
// reduce to the reference's argmax (largest FP64 dot, lowest index among ties); the float prefilter is
// the one of pk_common.cuh::support with the threshold taken over the whole hull.
template <int G> __device__ __forceinline__ d3 eg_support(const ShapeView &s, d3 d, int sub, unsigned pmask)
{
    constexpr int NSUB = G / 2;
    if (s.kind != KIND_HULL) return support(s, d);
    const d3 l = rotate(conjugate(s.q), d);
    const double *v = s.verts;
    uint32_t best = 0;
    double best_dot = 0.0;
    bool have = false;
    if (s.nverts <= HULL_PREFILTER_MIN)
    {
        for (uint32_t i = sub; i < s.nverts; i += NSUB)
        {
            double t = (v[3 * i] * l.x + v[3 * i + 1] * l.y) + v[3 * i + 2] * l.z;
            if (!have || t > best_dot)
            {
                best_dot = t;
                best = i;
                have = true;
            }
        }
    }
    else
    {
        const float lx = static_cast<float>(l.x), ly = static_cast<float>(l.y), lz = static_cast<float>(l.z);
        const float E = 1e-6f * s.hull_r * (fabsf(lx) + fabsf(ly) + fabsf(lz)) + 1e-37f;
        const float4 *__restrict__ vf = s.vf;
        float fm = -3.4e38f;
        for (uint32_t i = sub; i < s.nverts; i += NSUB)
        {
            float4 w = __ldg(vf + i);
            fm = fmaxf(fm, fmaf(w.x, lx, fmaf(w.y, ly, w.z * lz)));
        }
#pragma unroll
        for (int off = 2; off < G; off <<= 1) fm = fmaxf(fm, __shfl_xor_sync(pmask, fm, off));
        const float thr = fm - 2.0f * E;
        for (uint32_t i = sub; i < s.nverts; i += NSUB)
        {
            float4 w = __ldg(vf + i);
            if (fmaf(w.x, lx, fmaf(w.y, ly, w.z * lz)) >= thr)
            {
                double t = (v[3 * i] * l.x + v[3 * i + 1] * l.y) + v[3 * i + 2] * l.z;
                if (!have || t > best_dot)
                {
                    best_dot = t;
                    best = i;
                    have = true;
                }
            }
        }
    }
#pragma unroll
    for (int off = 2; off < G; off <<= 1)
    {
        const double od = __shfl_xor_sync(pmask, best_dot, off);
        const uint32_t oi = __shfl_xor_sync(pmask, best, off);
        const int oh = __shfl_xor_sync(pmask, have ? 1 : 0, off);
        if (oh && (!have || od > best_dot || (od == best_dot && oi < best)))
        {
            best_dot = od;
            best = oi;
            have = true;
        }
    }
    d3 bv{v[3 * best], v[3 * best + 1], v[3 * best + 2]};
    return rotate(s.q, bv) + s.p;
}

// collision.cpp:424-454.  u, v, w are computed by every lane; lane 0 blends the pa's and writes key,
// normal, world_a, depth; lane 1 blends the pb's and writes world_b.
template <class Poly>
__device__ __forceinline__ void eg_write_result(const Poly &po, int gl, double4 nd, unsigned long long t, ContactRec *out, uint64_t key)
{
    const d3 n{nd.x, nd.y, nd.z};
    const int i0 = eg_v(t, 0), i1 = eg_v(t, 1), i2 = eg_v(t, 2);
    const d3 p0 = po.vpos(i0), p1 = po.vpos(i1), p2 = po.vpos(i2);
    const d3 pm = n * nd.w;
    const d3 v0 = p1 - p0, v1 = p2 - p0, v2 = pm - p0;
    const double d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    const double denom = d00 * d11 - d01 * d01;
    const double v = (d11 * d20 - d01 * d21) / denom;
    const double w = (d00 * d21 - d01 * d20) / denom;
    const double u = 1.0 - v - w;
    if (gl < 2)
    {
        const double *q = po.g_vab + 3 * gl;
        const d3 x0{q[6 * i0], q[6 * i0 + 1], q[6 * i0 + 2]};
        const d3 x1{q[6 * i1], q[6 * i1 + 1], q[6 * i1 + 2]};
        const d3 x2{q[6 * i2], q[6 * i2 + 1], q[6 * i2 + 2]};
        const d3 ww = (u * x0 + v * x1) + w * x2;
        if (gl == 0)
        {
            out->key = key;
            out->normal[0] = -n.x;
            out->normal[1] = -n.y;
            out->normal[2] = -n.z;
            out->world_a[0] = ww.x;
            out->world_a[1] = ww.y;
            out->world_a[2] = ww.z;
            out->depth = nd.w;
        }
        else
        {
            out->world_b[0] = ww.x;
            out->world_b[1] = ww.y;
            out->world_b[2] = ww.z;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K7b (group form).  Same inputs and outputs as epa_kernel, plus the fallback list.
// ---------------------------------------------------------------------------------------------
// groups per block: as many polytopes as fit the 227 KB of one SM, in whole warps
template <int G, int F_S, int H_S, int V_S> constexpr int eg_groups_per_block()
{
    return static_cast<int>(232448 / sizeof(EgSmem<G, F_S, H_S, V_S>)) / (32 / G) * (32 / G);
}

template <int G, int F_S, int H_S, int V_S>
__global__ void __launch_bounds__(eg_groups_per_block<G, F_S, H_S, V_S>() * G, 1)
epa_group_kernel(BodyArrays bodies, const uint64_t *__restrict__ keys, const uint32_t *__restrict__ pair_a,
                 const uint32_t *__restrict__ pair_b, const SimplexRec *__restrict__ simplices,
                 const unsigned long long *__restrict__ hit_count_ptr, uint64_t hit_capacity,
                 const uint32_t *__restrict__ out_index, const uint32_t *__restrict__ order, ContactRec *__restrict__ contacts,
                 uint8_t *__restrict__ valid, unsigned char *__restrict__ spill, unsigned long long *__restrict__ cursor,
                 unsigned long long *__restrict__ counters /* [0]=valid contacts */, uint32_t *__restrict__ fallback_list,
                 unsigned long long *__restrict__ fallback_count)
{
    static_assert(G == 4 || G == 8, "group of 4 or 8 lanes");
    using SM = EgSmem<G, F_S, H_S, V_S>;
    using SP = EgSpillLayout<F_S, H_S, V_S>;
    static_assert(sizeof(SM) % 128 == SM::WANT, "group stride must spread the groups of a warp over the banks");
    extern __shared__ __align__(16) unsigned char eg_smem_raw[];

    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const int gbase = lane & ~(G - 1);
    const unsigned gmask = ((1u << G) - 1u) << gbase;
    const int parity = gl & 1;
    const int sub = gl >> 1;
    const unsigned pmask = gmask & (parity ? 0xAAAAAAAAu : 0x55555555u);
    const int group_in_block = threadIdx.x / G;
    const uint64_t group_global = static_cast<uint64_t>(blockIdx.x) * (blockDim.x / G) + group_in_block;

    EgPoly<G, F_S, H_S, V_S> po;
    po.s = reinterpret_cast<SM *>(eg_smem_raw + static_cast<size_t>(group_in_block) * sizeof(SM));
    {
        unsigned char *base = spill + group_global * SP::BYTES;
        po.g_plane = reinterpret_cast<double *>(base + SP::PLANE);
        po.g_topo = reinterpret_cast<unsigned long long *>(base + SP::TOPO);
        po.g_vpos = reinterpret_cast<double *>(base + SP::VPOS);
        po.g_vab = reinterpret_cast<double *>(base + SP::VAB);
    }
    for (int i = gl; i < 72; i += G) po.s->edge_of_start()[i] = po.s->edge_of_end()[i] = 0xFF;
    __syncwarp(gmask);

    unsigned long long nhits = *hit_count_ptr;
    if (nhits > hit_capacity) nhits = hit_capacity;

    bool active = false;
    int nverts = 0, nserial = 0, heap_size = 0, iter = 0;
    uint32_t out_slot = 0, cur_sidx = 0;
    uint64_t key = 0;
    unsigned long long fm0 = 0, fm1 = 0, fm2 = 0; // free face slots
    unsigned long long n_valid = 0;
    ShapeView mine{};

    auto to_fallback = [&]()
    {
        if (gl == 0)
        {
            unsigned long long i = atomicAdd(fallback_count, 1ull);
            if (i < hit_capacity) fallback_list[i] = cur_sidx;
        }
        active = false;
    };
    auto free_slot = [&](int f)
    {
        if (f < 64)
            fm0 |= 1ull << f;
        else if (f < 128)
            fm1 |= 1ull << (f - 64);
        else
            fm2 |= 1ull << (f - 128);
    };

    bool done = false;
    for (;;)
    {
        // The groups of a warp run the same loop on different pairs.  Without this barrier a group that
        // `continue`s early never rejoins the others and the warp ends up issuing every group's
        // instructions separately (measured: 2.2x slower than thread-per-pair).  With it the warp walks
        // through the phases of an iteration together.
        __syncwarp();
        if (__ballot_sync(0xFFFFFFFFu, !done) == 0u) break;
        if (done) continue;
        if (!active)
        {
            // ---- next pair of the hit list (order[] groups them by cost class) -----------------
            unsigned long long slot = 0;
            if (gl == 0) slot = atomicAdd(cursor, 1ull);
            slot = __shfl_sync(gmask, slot, gbase);
            if (slot >= nhits)
            {
                done = true;
                continue;
            }
            cur_sidx = order[slot];
            const SimplexRec *r = simplices + cur_sidx;
            const uint32_t rn = r->n & 0xFFu;
            const uint32_t pair = r->pair;
            if (rn != 4u)
            {
                to_fallback(); // pad_simplex path (collision.cpp:191-248): rare, left to epa_kernel
                continue;
            }
            uint32_t ia, ib;
            if (keys)
            {
                key = keys[pair];
                ia = static_cast<uint32_t>(key >> 32);
                ib = static_cast<uint32_t>(key & 0xFFFFFFFFu);
            }
            else
            {
                ia = pair_a[pair];
                ib = pair_b[pair];
                key = (static_cast<uint64_t>(ia) << 32) | ib;
            }
            out_slot = out_index[pair];
            mine = load_shape(bodies, parity ? ib : ia);
            if (gl < 4)
            {
                const double2 *q = reinterpret_cast<const double2 *>(&r->v[gl][0]);
                double2 a = q[0], b = q[1], c = q[2];
                d3 pa{a.x, a.y, b.x}, pb{b.y, c.x, c.y};
                po.set_vpos(gl, pa - pb);
                double2 *o = reinterpret_cast<double2 *>(po.g_vab + 6 * gl);
                o[0] = a;
                o[1] = b;
                o[2] = c;
            }
            __syncwarp(gmask);
            // build_initial_tetrahedron (collision.cpp:355-388): faces (0,1,2|3) (0,2,3|1) (0,3,1|2) (1,3,2|0)
            uint32_t myv = 0;
            double mydist = 0.0;
            if (gl < 4)
            {
                const int fi = (gl == 3) ? 1 : 0;
                const int fj = (gl == 0) ? 1 : (gl == 1 ? 2 : 3);
                const int fk = (gl == 0) ? 2 : (gl == 1 ? 3 : (gl == 2 ? 1 : 2));
                const int fo = (gl == 0) ? 3 : (gl == 1 ? 1 : (gl == 2 ? 2 : 0));
                d3 n;
                const bool flip = epa_face_plane(po.vpos(fi), po.vpos(fj), po.vpos(fk), true, po.vpos(fo), n, mydist);
                po.set_plane(gl, n, mydist);
                myv = static_cast<uint32_t>(fi) | (static_cast<uint32_t>(flip ? fk : fj) << 8) | (static_cast<uint32_t>(flip ? fj : fk) << 16);
            }
            uint32_t tv[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) tv[f] = __shfl_sync(gmask, myv, gbase + f);
            if (gl < 4)
            {
                // brute-force adjacency: an undirected tetrahedron edge belongs to exactly two faces, so a
                // directed edge has at most one reversed partner and the reference's i<j loop order is moot
                unsigned long long t = myv;
#pragma unroll
                for (int e1 = 0; e1 < 3; ++e1)
                {
                    const uint32_t u1 = (myv >> (8 * e1)) & 0xFFu, v1 = (myv >> (8 * ((e1 + 1) % 3))) & 0xFFu;
                    unsigned long long adj = 0xFFull;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        if (j == gl) continue;
#pragma unroll
                        for (int e2 = 0; e2 < 3; ++e2)
                        {
                            const uint32_t u2 = (tv[j] >> (8 * e2)) & 0xFFu, v2 = (tv[j] >> (8 * ((e2 + 1) % 3))) & 0xFFu;
                            if (u1 == v2 && v1 == u2) adj = static_cast<unsigned long long>(j);
                        }
                    }
                    t |= adj << (24 + 8 * e1);
                }
                t |= static_cast<unsigned long long>(gl) << 48;
                po.set_topo(gl, t);
            }
            heap_size = 0;
#pragma unroll
            for (int f = 0; f < 4; ++f)
            {
                const double df = __shfl_sync(gmask, mydist, gbase + f);
                eg_heap_push(po.s, heap_size, df, (static_cast<uint32_t>(f) << 16) | static_cast<uint32_t>(f));
            }
            fm0 = ~0xFull;
            fm1 = ~0ull;
            fm2 = (1ull << (EG_MAX_SLOTS - 128)) - 1ull;
            nverts = 4;
            nserial = 4;
            iter = 0;
            active = true;
            __syncwarp(gmask);
        }

        // ---- one EPA iteration, or the post-loop "best guess" when iter == 64 -----------------
        int min_slot = -1;
        unsigned long long mt = 0;
        while (heap_size > 0) // pop_face(): skip obsolete entries (collision.cpp:397-408)
        {
            const uint32_t id = eg_heap_pop(po.s, heap_size);
            const int sl = static_cast<int>(id & 0xFFu);
            const unsigned long long t = po.topo(sl);
            if (eg_serial(t) == (id >> 16))
            {
                min_slot = sl;
                mt = t;
                break;
            }
        }
        if (min_slot < 0)
        {
            if (gl == 0) valid[out_slot] = 0; // heap exhausted → nullopt (collision.cpp:459,502)
            active = false;
            continue;
        }
        const double4 mf = po.plane(min_slot);
        if (iter >= 64)
        {
            eg_write_result(po, gl, mf, mt, contacts + out_slot, key); // best guess (collision.cpp:500-503)
            if (gl == 0)
            {
                valid[out_slot] = 1;
                ++n_valid;
            }
            active = false;
            continue;
        }
        ++iter;
        const d3 mn{mf.x, mf.y, mf.z};
        const d3 sup = eg_support<G>(mine, parity ? -mn : mn, sub, pmask); // collision.h:41-49
        const d3 oth = eg_shfl_xor(gmask, sup, 1);
        const d3 p = parity ? (oth - sup) : (sup - oth);
        if (dot(mn, p) - mf.w < 1e-6)
        {
            eg_write_result(po, gl, mf, mt, contacts + out_slot, key); // converged (collision.cpp:465-466)
            if (gl == 0)
            {
                valid[out_slot] = 1;
                ++n_valid;
            }
            active = false;
            continue;
        }

        // find_silhouette (collision.cpp:315-353): LIFO flood fill, edge order preserved.  Lanes 0-2 test
        // the three neighbours of the current face; the bookkeeping is done by every lane.
        bool bad = false;
        int nh = 0;
        {
            po.set_topo(min_slot, mt | EG_DEAD);
            free_slot(min_slot);
            unsigned long long stack = 0;
            int depth = 0;
            unsigned long long cur = mt;
            for (;;)
            {
                const int a_my = (gl < 3) ? eg_adj(cur, gl) : 0xFF;
                bool live = false, vis = false;
                if (a_my != 0xFF)
                {
                    live = eg_serial(po.topo(a_my)) != 0xFFFFu;
                    if (live)
                    {
                        const double4 q = po.plane(a_my);
                        vis = dot(d3{q.x, q.y, q.z}, p) > q.w + 1e-6;
                    }
                }
                const unsigned bl = __ballot_sync(gmask, live) >> gbase;
                const unsigned bv = __ballot_sync(gmask, vis) >> gbase;
#pragma unroll
                for (int i = 0; i < 3; ++i)
                {
                    if (!((bl >> i) & 1u)) continue;
                    const int ai = eg_adj(cur, i);
                    // a neighbour reached through two edges of this face: the second visit sees the
                    // obsolete flag set by the first one
                    bool dup = false;
#pragma unroll
                    for (int j = 0; j < i; ++j)
                        if (((bl >> j) & 1u) && ((bv >> j) & 1u) && eg_adj(cur, j) == ai) dup = true;
                    if (dup) continue;
                    if ((bv >> i) & 1u)
                    {
                        po.set_topo(ai, po.topo(ai) | EG_DEAD);
                        free_slot(ai);
                        if (depth < 8)
                        {
                            stack = (stack << 8) | static_cast<unsigned long long>(ai);
                            ++depth;
                        }
                        else
                            bad = true;
                    }
                    else
                    {
                        if (nh < EG_MAX_HORIZON)
                        {
                            po.s->hz_start[nh] = static_cast<uint8_t>(eg_v(cur, i));
                            po.s->hz_end[nh] = static_cast<uint8_t>(eg_v(cur, (i + 1) % 3));
                            po.s->hz_adj[nh] = static_cast<uint8_t>(ai);
                            ++nh;
                        }
                        else
                            bad = true;
                    }
                }
                if (depth == 0) break;
                cur = po.topo(static_cast<int>(stack & 0xFFull));
                stack >>= 8;
                --depth;
            }
        }
        if (nh == 0 && !bad)
        {
            iter = 64; // empty horizon → best remaining face (collision.cpp:469,500-503)
            continue;
        }
        if (bad || heap_size + nh > H_S || nh < 3)
        {
            to_fallback();
            continue;
        }
        // the new vertex
        const int p_idx = nverts++;
        po.set_vpos(p_idx, p);
        if (gl < 2)
        {
            double *q = po.g_vab + 6 * p_idx + 3 * gl;
            q[0] = sup.x;
            q[1] = sup.y;
            q[2] = sup.z;
        }
        // slots for the new faces, lowest free first (keeps the live set inside the shared-memory part)
        if (__popcll(fm0) >= nh)
        {
            for (int e = 0; e < nh; ++e)
            {
                po.s->hz_slot[e] = static_cast<uint8_t>(__ffsll(static_cast<long long>(fm0)) - 1);
                fm0 &= fm0 - 1;
            }
        }
        else
        {
            for (int e = 0; e < nh; ++e)
            {
                int sl = 0;
                if (fm0)
                {
                    sl = __ffsll(static_cast<long long>(fm0)) - 1;
                    fm0 &= fm0 - 1;
                }
                else if (fm1)
                {
                    sl = 64 + __ffsll(static_cast<long long>(fm1)) - 1;
                    fm1 &= fm1 - 1;
                }
                else if (fm2)
                {
                    sl = 128 + __ffsll(static_cast<long long>(fm2)) - 1;
                    fm2 &= fm2 - 1;
                }
                else
                    bad = true;
                po.s->hz_slot[e] = static_cast<uint8_t>(sl);
            }
        }
        if (bad)
        {
            to_fallback();
            continue;
        }
        // new faces (start, end, p_idx), no orientation flip (collision.cpp:475-482): one per lane
        for (int base = 0; base < nh; base += G)
        {
            const int e = base + gl;
            if (e < nh)
            {
                const int st = po.s->hz_start[e], en = po.s->hz_end[e], ad = po.s->hz_adj[e], sl = po.s->hz_slot[e];
                d3 n;
                double dist;
                epa_face_plane(po.vpos(st), po.vpos(en), p, false, d3{0, 0, 0}, n, dist);
                po.set_plane(sl, n, dist);
                // link_faces(f, adj_face, start, end): on the old face the shared edge starts at `end`
                const unsigned long long tb = po.topo(ad);
                const int e2 = (eg_v(tb, 0) == en) ? 0 : (eg_v(tb, 1) == en ? 1 : 2);
                if (eg_v(tb, e2) != en) bad = true; // unmatched link: slot recycling is no longer safe
                po.set_adj(ad, e2, sl);
                po.s->edge_of_start()[st] = static_cast<uint8_t>(e);
                po.s->edge_of_end()[en] = static_cast<uint8_t>(e);
            }
        }
        __syncwarp(gmask);
        // ring links among the new faces (collision.cpp:484-497) through the vertex→edge tables; valid for
        // a proper horizon (every vertex starts at most one edge and ends at most one, no self loop, no
        // 2-cycle) — anything else goes to the fallback, which runs the reference's i<j loop
        for (int base = 0; base < nh; base += G)
        {
            const int e = base + gl;
            if (e < nh)
            {
                const int st = po.s->hz_start[e], en = po.s->hz_end[e], ad = po.s->hz_adj[e], sl = po.s->hz_slot[e];
                if (st == en || po.s->edge_of_start()[st] != e || po.s->edge_of_end()[en] != e) bad = true;
                const int j = po.s->edge_of_start()[en];
                const int k = po.s->edge_of_end()[st];
                if (j != 0xFF && po.s->edge_of_start()[po.s->hz_end[j]] == e) bad = true;
                const int a1 = (j == 0xFF) ? 0xFF : po.s->hz_slot[j];
                const int a2 = (k == 0xFF) ? 0xFF : po.s->hz_slot[k];
                const unsigned long long t = static_cast<unsigned long long>(st) | (static_cast<unsigned long long>(en) << 8) |
                                             (static_cast<unsigned long long>(p_idx) << 16) | (static_cast<unsigned long long>(ad) << 24) |
                                             (static_cast<unsigned long long>(a1) << 32) | (static_cast<unsigned long long>(a2) << 40) |
                                             (static_cast<unsigned long long>(nserial + e) << 48);
                po.set_topo(sl, t);
            }
        }
        __syncwarp(gmask);
        for (int base = 0; base < nh; base += G)
        {
            const int e = base + gl;
            if (e < nh)
            {
                po.s->edge_of_start()[po.s->hz_start[e]] = 0xFF;
                po.s->edge_of_end()[po.s->hz_end[e]] = 0xFF;
            }
        }
        bad = __any_sync(gmask, bad);
        __syncwarp(gmask);
        if (bad)
        {
            to_fallback();
            continue;
        }
        for (int e = 0; e < nh; ++e) // push_face in horizon order
        {
            const int sl = po.s->hz_slot[e];
            eg_heap_push(po.s, heap_size, po.plane_dist(sl), (static_cast<uint32_t>(nserial + e) << 16) | static_cast<uint32_t>(sl));
        }
        nserial += nh;
    }
    if (gl == 0 && n_valid) atomicAdd(counters + 0, n_valid);
}

} // namespace pk
