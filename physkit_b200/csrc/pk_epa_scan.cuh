// pk_epa_scan.cuh — EPA (reference src/collision.cpp:251-509), one thread per pair, organised around
// the number of DEPENDENT memory round trips per iteration.
//
// ncu on epa_kernel (pk_narrowphase.cuh) at 1 M bodies: 71 % of stall samples are long_scoreboard, spread
// over the heap sifts (one HBM/L2 round trip per level below the shared-memory top), the flood fill (plane,
// then topology, per visited face) and the face-creation loop (vertex and neighbour loads per horizon
// edge): ≈24 dependent round trips per iteration, ≈1 µs each under load.  This kernel removes most of them:
//
//   * no heap.  The reference pops the live face with the smallest distance; which face that is depends
//     on the heap's internal order only when two live faces tie for the minimum.  A float key
//     (distance rounded down) per face slot lives in shared memory; pop_face is a scan of those keys
//     (no memory round trip), exact distances are consulted only when several slots share the minimal
//     key, and an exact tie hands the pair to epa_kernel, which restates the heap (≈10 % of the pairs;
//     zero-distance ties created by one iteration are resolved here, see below)
//   * face slots are recycled (a face made obsolete frees its slot), so the live polytope stays in the
//     first 2V−4 slots of a small, dense per-thread slab and the obsolete test is `key == +inf` in shared
//     memory instead of a bitset in local memory
//   * the flood fill loads plane and topology of the three neighbours together and prefetches the
//     neighbourhood of every face it pushes; the face loop prefetches all its operands first
//   * horizon and ring-link scratch live in shared memory, not in local memory
//
// Capacity is cut to what shared memory holds (96 face slots ⇒ 46 iterations, 16 horizon edges); pairs
// beyond it, padded simplices and improper horizons also go to epa_kernel through the fallback list.
// Results are bit-identical to epa_kernel.
#pragma once

#include "pk_narrowphase.cuh"

namespace pk
{

constexpr int ES_THREADS = 64;
constexpr int ES_SLOTS = 96;   // live faces = 2V − 4
constexpr int ES_VERTS = 50;   // 2·50 − 4 = 96
constexpr int ES_HORIZON = 16; // observed max 10
constexpr int ES_STACK = 8;    // observed max 4
constexpr size_t ES_SLAB_BYTES = static_cast<size_t>(ES_SLOTS) * (32 + 8) + static_cast<size_t>(ES_VERTS) * (32 + 48);

struct EsSmem
{
    float key[ES_SLOTS][ES_THREADS]; // float(distance) rounded down; +inf = free slot
    double f[2][10][ES_THREADS];     // shape views: p xyz, h xyz, q xyzw
    const double *verts[2][ES_THREADS];
    const float4 *vf[2][ES_THREADS];
    float hull_r[2][ES_THREADS];
    int kind[2][ES_THREADS];
    uint32_t nverts[2][ES_THREADS];
    uint32_t hz[ES_HORIZON][ES_THREADS];   // start | end << 8 | adjacent slot << 16 | new slot << 24
    uint16_t ring[ES_HORIZON][ES_THREADS]; // successor slot | predecessor slot << 8 (0xFF = none)
    uint8_t edge_of_start[ES_VERTS][ES_THREADS];
};
static_assert(sizeof(EsSmem) <= 48 * 1024, "EsSmem must fit static shared memory");

struct EsSlab
{
    double *plane;            // normal xyz, distance: one 32-byte sector per slot
    unsigned long long *topo; // bytes 0-2 vertices, 3-5 adjacent slots (0xFF = none)
    double *vpos;             // p = pa − pb, padded to 32 bytes
    double *vab;              // pa xyz, pb xyz
    __device__ __forceinline__ explicit EsSlab(unsigned char *base)
    {
        plane = reinterpret_cast<double *>(base);
        base += static_cast<size_t>(ES_SLOTS) * 32;
        topo = reinterpret_cast<unsigned long long *>(base);
        base += static_cast<size_t>(ES_SLOTS) * 8;
        vpos = reinterpret_cast<double *>(base);
        base += static_cast<size_t>(ES_VERTS) * 32;
        vab = reinterpret_cast<double *>(base);
    }
    __device__ __forceinline__ double4 load_plane(int f) const
    {
        const double2 *q = reinterpret_cast<const double2 *>(plane + 4 * f);
        double2 a = q[0], b = q[1];
        return make_double4(a.x, a.y, b.x, b.y);
    }
    __device__ __forceinline__ void store_plane(int f, d3 n, double dist) const
    {
        double2 *q = reinterpret_cast<double2 *>(plane + 4 * f);
        q[0] = make_double2(n.x, n.y);
        q[1] = make_double2(n.z, dist);
    }
    __device__ __forceinline__ d3 vp(int i) const
    {
        const double2 *q = reinterpret_cast<const double2 *>(vpos + 4 * i);
        double2 a = q[0], b = q[1];
        return {a.x, a.y, b.x};
    }
    __device__ __forceinline__ void set_vert(int i, const SupportPt &s, d3 p) const
    {
        double2 *q = reinterpret_cast<double2 *>(vpos + 4 * i);
        q[0] = make_double2(p.x, p.y);
        q[1] = make_double2(p.z, 0.0);
        double2 *v = reinterpret_cast<double2 *>(vab + 6 * i);
        v[0] = make_double2(s.pa.x, s.pa.y);
        v[1] = make_double2(s.pa.z, s.pb.x);
        v[2] = make_double2(s.pb.y, s.pb.z);
    }
    __device__ __forceinline__ void set_adj(int f, int e, int to) const
    {
        reinterpret_cast<uint8_t *>(topo + f)[3 + e] = static_cast<uint8_t>(to);
    }
};

__device__ __forceinline__ int es_v(unsigned long long t, int i) { return static_cast<int>((t >> (8 * i)) & 0xFFull); }
__device__ __forceinline__ int es_adj(unsigned long long t, int i) { return static_cast<int>((t >> (24 + 8 * i)) & 0xFFull); }
__device__ __forceinline__ void es_prefetch(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ float es_inf() { return __int_as_float(0x7F800000); }

__device__ __forceinline__ void es_put_shape(EsSmem &sm, int which, const ShapeView &v)
{
    const int t = threadIdx.x;
    sm.f[which][0][t] = v.p.x; sm.f[which][1][t] = v.p.y; sm.f[which][2][t] = v.p.z;
    sm.f[which][3][t] = v.h.x; sm.f[which][4][t] = v.h.y; sm.f[which][5][t] = v.h.z;
    sm.f[which][6][t] = v.q.x; sm.f[which][7][t] = v.q.y; sm.f[which][8][t] = v.q.z; sm.f[which][9][t] = v.q.w;
    sm.verts[which][t] = v.verts;
    sm.vf[which][t] = v.vf;
    sm.hull_r[which][t] = v.hull_r;
    sm.kind[which][t] = v.kind;
    sm.nverts[which][t] = v.nverts;
}
__device__ __forceinline__ ShapeView es_get_shape(const EsSmem &sm, int which)
{
    const int t = threadIdx.x;
    ShapeView v;
    v.p = {sm.f[which][0][t], sm.f[which][1][t], sm.f[which][2][t]};
    v.h = {sm.f[which][3][t], sm.f[which][4][t], sm.f[which][5][t]};
    v.q = {sm.f[which][6][t], sm.f[which][7][t], sm.f[which][8][t], sm.f[which][9][t]};
    v.verts = sm.verts[which][t];
    v.vf = sm.vf[which][t];
    v.hull_r = sm.hull_r[which][t];
    v.kind = sm.kind[which][t];
    v.nverts = sm.nverts[which][t];
    return v;
}

// collision.cpp:424-454
__device__ __forceinline__ void es_write_result(const EsSlab &sl, double4 nd, unsigned long long t, ContactRec *out, uint64_t key)
{
    d3 n{nd.x, nd.y, nd.z};
    const double2 *q0 = reinterpret_cast<const double2 *>(sl.vab + 6 * es_v(t, 0));
    const double2 *q1 = reinterpret_cast<const double2 *>(sl.vab + 6 * es_v(t, 1));
    const double2 *q2 = reinterpret_cast<const double2 *>(sl.vab + 6 * es_v(t, 2));
    double2 x0 = q0[0], x1 = q0[1], x2 = q0[2], y0 = q1[0], y1 = q1[1], y2 = q1[2], z0 = q2[0], z1 = q2[1], z2 = q2[2];
    d3 a0{x0.x, x0.y, x1.x}, b0{x1.y, x2.x, x2.y};
    d3 a1{y0.x, y0.y, y1.x}, b1{y1.y, y2.x, y2.y};
    d3 a2{z0.x, z0.y, z1.x}, b2{z1.y, z2.x, z2.y};
    d3 p0 = a0 - b0, p1 = a1 - b1, p2 = a2 - b2;
    d3 pm = n * nd.w;
    d3 v0 = p1 - p0, v1 = p2 - p0, v2 = pm - p0;
    double d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    double denom = d00 * d11 - d01 * d01;
    double v = (d11 * d20 - d01 * d21) / denom;
    double w = (d00 * d21 - d01 * d20) / denom;
    double u = 1.0 - v - w;
    d3 wa = (u * a0 + v * a1) + w * a2;
    d3 wb = (u * b0 + v * b1) + w * b2;
    out->key = key;
    out->normal[0] = -n.x;
    out->normal[1] = -n.y;
    out->normal[2] = -n.z;
    out->world_a[0] = wa.x;
    out->world_a[1] = wa.y;
    out->world_a[2] = wa.z;
    out->world_b[0] = wb.x;
    out->world_b[1] = wb.y;
    out->world_b[2] = wb.z;
    out->depth = nd.w;
}

#ifndef PK_ES_MIN_BLOCKS
#define PK_ES_MIN_BLOCKS 4
#endif
#ifndef PK_ES_FETCH_MIN
#define PK_ES_FETCH_MIN 6
#endif

__global__ void __launch_bounds__(ES_THREADS, PK_ES_MIN_BLOCKS)
epa_scan_kernel(BodyArrays bodies, const uint64_t *__restrict__ keys, const uint32_t *__restrict__ pair_a,
                const uint32_t *__restrict__ pair_b, const SimplexRec *__restrict__ simplices,
                const unsigned long long *__restrict__ hit_count_ptr, uint64_t hit_capacity,
                const uint32_t *__restrict__ out_index, const uint32_t *__restrict__ order, ContactRec *__restrict__ contacts,
                uint8_t *__restrict__ valid, unsigned char *__restrict__ slabs, unsigned long long *__restrict__ cursor,
                unsigned long long *__restrict__ counters /* [0]=valid contacts */, uint32_t *__restrict__ fallback_list,
                unsigned long long *__restrict__ fallback_count)
{
    __shared__ EsSmem shm;
    const int t = threadIdx.x;
    const uint64_t tid = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    const EsSlab sl(slabs + tid * ES_SLAB_BYTES);
    unsigned long long nhits = *hit_count_ptr;
    if (nhits > hit_capacity) nhits = hit_capacity;
    const float INF = es_inf();

    bool active = false, done = false;
    int nverts = 0, iter = 0, hi = 0; // hi: slots [0, hi) have been used by the current polytope
    unsigned long long fm0 = 0;       // free slots 0-63
    uint32_t fm1 = 0;                 // free slots 64-95
    uint32_t out_slot = 0, cur_sidx = 0;
    uint64_t key = 0;
    // zero-distance ties (see pop): all faces created before the last batch were strictly farther than 0
    bool older_positive = true, batch_positive = true;
    int batch_first = 0, batch_n = 0; // hz[] entries of the last batch are still valid when batch_n > 0
    unsigned long long n_valid = 0;
    for (int s = 0; s < ES_SLOTS; ++s) shm.key[s][t] = INF;
    for (int v = 0; v < ES_VERTS; ++v) shm.edge_of_start[v][t] = 0xFF;
    constexpr unsigned FULL = 0xFFFFFFFFu;

    auto to_fallback = [&]()
    {
        unsigned long long i = atomicAdd(fallback_count, 1ull);
        if (i < hit_capacity) fallback_list[i] = cur_sidx;
        active = false;
    };
    auto kill_slot = [&](int f)
    {
        shm.key[f][t] = INF;
        if (f < 64)
            fm0 |= 1ull << f;
        else
            fm1 |= 1u << (f - 64);
    };

    for (;;)
    {
        const unsigned m_active = __ballot_sync(FULL, active);
        const unsigned m_idle = __ballot_sync(FULL, !active && !done);
        if (m_active == 0 && m_idle == 0) break;
        if (!active && !done && (__popc(m_idle) >= PK_ES_FETCH_MIN || m_active == 0))
        {
            unsigned long long slot = atomicAdd(cursor, 1ull);
            if (slot >= nhits)
                done = true;
            else
            {
                cur_sidx = order[slot];
                const SimplexRec *r = simplices + cur_sidx;
                const uint32_t pair = r->pair;
                if ((r->n & 0xFFu) != 4u)
                    to_fallback(); // pad_simplex path (collision.cpp:191-248): rare, left to epa_kernel
                else
                {
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
                    {
                        ShapeView A = load_shape(bodies, ia);
                        ShapeView B = load_shape(bodies, ib);
                        es_put_shape(shm, 0, A);
                        es_put_shape(shm, 1, B);
                    }
                    d3 pv[4];
                    for (int i = 0; i < 4; ++i)
                    {
                        SupportPt s;
                        s.pa = d3{r->v[i][0], r->v[i][1], r->v[i][2]};
                        s.pb = d3{r->v[i][3], r->v[i][4], r->v[i][5]};
                        pv[i] = P(s);
                        sl.set_vert(i, s, pv[i]);
                    }
                    for (int s = 4; s < hi; ++s) shm.key[s][t] = INF; // leftovers of the previous polytope
                    // build_initial_tetrahedron (collision.cpp:355-388): faces (0,1,2|3) (0,2,3|1) (0,3,1|2) (1,3,2|0)
                    const int fi[4] = {0, 0, 0, 1}, fj[4] = {1, 2, 3, 3}, fk[4] = {2, 3, 1, 2}, fo[4] = {3, 1, 2, 0};
                    uint8_t tv[4][3], ta[4][3];
                    bool bad = false;
                    batch_positive = true;
                    for (int f = 0; f < 4; ++f)
                    {
                        d3 n;
                        double dist;
                        bool flip = epa_face_plane(pv[fi[f]], pv[fj[f]], pv[fk[f]], true, pv[fo[f]], n, dist);
                        tv[f][0] = static_cast<uint8_t>(fi[f]);
                        tv[f][1] = static_cast<uint8_t>(flip ? fk[f] : fj[f]);
                        tv[f][2] = static_cast<uint8_t>(flip ? fj[f] : fk[f]);
                        ta[f][0] = ta[f][1] = ta[f][2] = 0xFF;
                        sl.store_plane(f, n, dist);
                        if (!(fabs(dist) < 1e30)) bad = true; // NaN / inf: the key order would not be the heap's
                        if (!(dist > 0.0)) batch_positive = false;
                        shm.key[f][t] = __double2float_rd(dist);
                        // horizon-order record of this batch (used only to resolve zero-distance ties)
                        shm.hz[f][t] = static_cast<uint32_t>(f) << 24;
                    }
                    for (int i = 0; i < 4; ++i)
                        for (int j = i + 1; j < 4; ++j)
                            for (int e1 = 0; e1 < 3; ++e1)
                            {
                                uint8_t u1 = tv[i][e1], v1 = tv[i][(e1 + 1) % 3];
                                for (int e2 = 0; e2 < 3; ++e2)
                                {
                                    uint8_t u2 = tv[j][e2], v2 = tv[j][(e2 + 1) % 3];
                                    if (u1 == v2 && v1 == u2)
                                    {
                                        ta[i][e1] = static_cast<uint8_t>(j);
                                        ta[j][e2] = static_cast<uint8_t>(i);
                                    }
                                }
                            }
                    for (int f = 0; f < 4; ++f)
                    {
                        unsigned long long w = 0;
                        for (int k = 0; k < 3; ++k)
                            w |= (static_cast<unsigned long long>(tv[f][k]) << (8 * k)) | (static_cast<unsigned long long>(ta[f][k]) << (24 + 8 * k));
                        sl.topo[f] = w;
                    }
                    fm0 = ~0xFull;
                    fm1 = 0xFFFFFFFFu;
                    hi = 4;
                    nverts = 4;
                    iter = 0;
                    older_positive = true;
                    batch_first = 0;
                    batch_n = 4;
                    active = true;
                    if (bad) to_fallback();
                }
            }
        }
        if (!active) continue;

        // ---- pop_face (collision.cpp:397-408) without a heap: the live face with the smallest distance ----
        int min_face = -1;
        {
            float m = INF;
            int cnt = 0;
            for (int s = 0; s < hi; ++s)
            {
                const float k = shm.key[s][t];
                if (k < m)
                {
                    m = k;
                    min_face = s;
                    cnt = 1;
                }
                else if (k == m)
                    ++cnt;
            }
            if (min_face >= 0 && cnt > 1)
            {
                // several live faces share the minimal float key: compare the exact distances
                double best = sl.plane[4 * min_face + 3];
                bool tie = false;
                for (int s = min_face + 1; s < hi; ++s)
                {
                    if (shm.key[s][t] != m) continue;
                    const double d = sl.plane[4 * s + 3];
                    if (d < best)
                    {
                        best = d;
                        min_face = s;
                        tie = false;
                    }
                    else if (d == best)
                        tie = true;
                }
                if (tie)
                {
                    // Which of two equidistant faces std::pop_heap delivers depends on the heap's history,
                    // with one provable exception: if every entry ever pushed before the last batch was
                    // strictly positive and the minimum is ±0, the first zero pushed by that batch sifted
                    // up to the root (all ancestors > 0) and no later zero passes it (__push_heap moves a
                    // parent down only if parent > value).  That is the degenerate-face case (normal 0,
                    // distance 0, collision.cpp:282-287) which ends EPA at this pop.
                    bool resolved = false;
                    if (best == 0.0 && older_positive && batch_n > 0)
                    {
                        for (int e = 0; e < batch_n && !resolved; ++e)
                        {
                            const int s = static_cast<int>(shm.hz[batch_first + e][t] >> 24);
                            if (shm.key[s][t] != INF && sl.plane[4 * s + 3] == 0.0)
                            {
                                min_face = s;
                                resolved = true;
                            }
                        }
                    }
                    if (!resolved)
                    {
                        to_fallback();
                        continue;
                    }
                }
            }
        }
        if (min_face < 0)
        {
            valid[out_slot] = 0; // heap exhausted → nullopt (collision.cpp:459,502)
            active = false;
            continue;
        }
        const double4 mf = sl.load_plane(min_face);
        const unsigned long long mt = sl.topo[min_face];
        if (iter >= 64)
        {
            es_write_result(sl, mf, mt, contacts + out_slot, key); // best guess (collision.cpp:500-503)
            valid[out_slot] = 1;
            ++n_valid;
            active = false;
            continue;
        }
        ++iter;
        const d3 mn{mf.x, mf.y, mf.z};
        SupportPt sp;
        {
            ShapeView A = es_get_shape(shm, 0);
            sp.pa = support(A, mn);
        }
        {
            ShapeView B = es_get_shape(shm, 1);
            sp.pb = support(B, -mn);
        }
        const d3 p = P(sp);
        if (dot(mn, p) - mf.w < 1e-6)
        {
            es_write_result(sl, mf, mt, contacts + out_slot, key); // converged (collision.cpp:465-466)
            valid[out_slot] = 1;
            ++n_valid;
            active = false;
            continue;
        }

        // ---- find_silhouette (collision.cpp:315-353): LIFO flood fill, edge order preserved ----
        bool bad = false;
        int nh = 0;
        {
            kill_slot(min_face);
            unsigned long long stack = 0;
            int depth = 0;
            unsigned long long cur = mt;
            for (;;)
            {
                double4 nf[3];
                unsigned long long nt[3];
                bool live[3];
#pragma unroll
                for (int i = 0; i < 3; ++i)
                {
                    const int a = es_adj(cur, i);
                    live[i] = a != 0xFF && shm.key[a][t] != INF;
                    if (live[i])
                    {
                        nf[i] = sl.load_plane(a);
                        nt[i] = sl.topo[a];
                    }
                }
#pragma unroll
                for (int i = 0; i < 3; ++i)
                {
                    if (!live[i]) continue;
                    const int a = es_adj(cur, i);
                    if (shm.key[a][t] == INF) continue; // reached through two edges of this face: first visit killed it
                    if (dot(d3{nf[i].x, nf[i].y, nf[i].z}, p) > nf[i].w + 1e-6)
                    {
                        kill_slot(a);
                        if (depth < ES_STACK)
                        {
                            stack = (stack << 8) | static_cast<unsigned long long>(a);
                            ++depth;
                            // its neighbourhood will be needed when it is popped
#pragma unroll
                            for (int k = 0; k < 3; ++k)
                            {
                                const int b = es_adj(nt[i], k);
                                if (b != 0xFF)
                                {
                                    es_prefetch(sl.plane + 4 * b);
                                    es_prefetch(sl.topo + b);
                                }
                            }
                        }
                        else
                            bad = true;
                    }
                    else
                    {
                        if (nh < ES_HORIZON)
                        {
                            shm.hz[nh][t] = static_cast<uint32_t>(es_v(cur, i)) | (static_cast<uint32_t>(es_v(cur, (i + 1) % 3)) << 8) |
                                            (static_cast<uint32_t>(a) << 16);
                            ++nh;
                        }
                        else
                            bad = true;
                    }
                }
                if (depth == 0) break;
                cur = sl.topo[static_cast<int>(stack & 0xFFull)];
                stack >>= 8;
                --depth;
            }
        }
        batch_n = 0;
        if (nh == 0 && !bad)
        {
            iter = 64; // empty horizon → best remaining face (collision.cpp:469,500-503)
            continue;
        }
        const int nfree = __popcll(fm0) + __popc(fm1);
        if (bad || nh < 3 || nfree < nh || nverts >= ES_VERTS)
        {
            to_fallback();
            continue;
        }
        // operands of the face loop: horizon vertices and the topology of the faces across the horizon
        for (int e = 0; e < nh; ++e)
        {
            const uint32_t h = shm.hz[e][t];
            es_prefetch(sl.vpos + 4 * (h & 0xFFu));
            es_prefetch(sl.topo + ((h >> 16) & 0xFFu));
        }
        sl.set_vert(nverts, sp, p);
        const int p_idx = nverts++;
        older_positive = older_positive && batch_positive;
        batch_positive = true;
        // new faces (start, end, p_idx), no orientation flip (collision.cpp:475-482), slots lowest free first
        unsigned long long end_seen = 0;
        for (int e = 0; e < nh; ++e)
        {
            const uint32_t h = shm.hz[e][t];
            const int st = static_cast<int>(h & 0xFFu), en = static_cast<int>((h >> 8) & 0xFFu), ad = static_cast<int>((h >> 16) & 0xFFu);
            int slot;
            if (fm0)
            {
                slot = __ffsll(static_cast<long long>(fm0)) - 1;
                fm0 &= fm0 - 1;
            }
            else
            {
                slot = 64 + __ffs(static_cast<int>(fm1)) - 1;
                fm1 &= fm1 - 1;
            }
            if (slot >= hi) hi = slot + 1;
            d3 n;
            double dist;
            epa_face_plane(sl.vp(st), sl.vp(en), p, false, d3{0, 0, 0}, n, dist);
            sl.store_plane(slot, n, dist);
            if (!(fabs(dist) < 1e30)) bad = true;
            if (!(dist > 0.0)) batch_positive = false;
            shm.key[slot][t] = __double2float_rd(dist);
            shm.hz[e][t] = h | (static_cast<uint32_t>(slot) << 24);
            shm.ring[e][t] = 0xFFFFu;
            // link_faces(f, adj_face, start, end): on the old face the shared edge starts at `end`
            const unsigned long long tb = sl.topo[ad];
            const int e2 = (es_v(tb, 0) == en) ? 0 : (es_v(tb, 1) == en ? 1 : 2);
            if (es_v(tb, e2) != en) bad = true; // unmatched link: slot recycling is no longer safe
            sl.set_adj(ad, e2, slot);
            // proper horizon: every vertex starts at most one edge and ends at most one, no self loop
            if (st == en || shm.edge_of_start[st][t] != 0xFF || ((end_seen >> en) & 1ull)) bad = true;
            shm.edge_of_start[st][t] = static_cast<uint8_t>(e);
            end_seen |= 1ull << en;
        }
        // ring links among the new faces (collision.cpp:484-497): face e = (start, end, p_idx) gets its
        // successor (the edge starting at `end`) on edge 1 and is that face's neighbour on edge 2
        for (int e = 0; e < nh; ++e)
        {
            const uint32_t h = shm.hz[e][t];
            const int j = shm.edge_of_start[(h >> 8) & 0xFFu][t];
            if (j == 0xFF) continue;
            const uint32_t hj = shm.hz[j][t];
            if (shm.edge_of_start[(hj >> 8) & 0xFFu][t] == e) bad = true; // 2-cycle: the reference links it one way only
            shm.ring[e][t] = static_cast<uint16_t>((shm.ring[e][t] & 0xFF00u) | (hj >> 24));
            shm.ring[j][t] = static_cast<uint16_t>((shm.ring[j][t] & 0x00FFu) | ((h >> 24) << 8));
        }
        for (int e = 0; e < nh; ++e)
        {
            const uint32_t h = shm.hz[e][t];
            const uint32_t rg = shm.ring[e][t];
            shm.edge_of_start[h & 0xFFu][t] = 0xFF;
            sl.topo[h >> 24] = static_cast<unsigned long long>(h & 0xFFFFu) | (static_cast<unsigned long long>(p_idx) << 16) |
                               (static_cast<unsigned long long>((h >> 16) & 0xFFu) << 24) |
                               (static_cast<unsigned long long>(rg & 0xFFu) << 32) | (static_cast<unsigned long long>(rg >> 8) << 40);
        }
        if (bad)
        {
            to_fallback();
            continue;
        }
        batch_first = 0;
        batch_n = nh;
    }
    if (n_valid) atomicAdd(counters + 0, n_valid);
}

} // namespace pk
