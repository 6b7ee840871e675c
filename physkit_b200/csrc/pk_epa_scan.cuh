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
// Shared memory holds keys for the first 96 face slots (46 iterations); the few polytopes that grow past
// that scan the remaining distances from their slab.  Padded simplices, improper horizons and horizons
// longer than ES_HORIZON edges also go to epa_kernel through the fallback list.
// Results are bit-identical to epa_kernel.
#pragma once

#include "pk_narrowphase.cuh"
// -DPK_ES_WHY: debug build that prints why the HEAP instance hands a pair back (which `bad` site fired)
#ifdef PK_ES_WHY
#include <cstdio>
#define PK_ES_WHY_TAG(b) why |= (b)
#else
#define PK_ES_WHY_TAG(b)
#endif

namespace pk
{

constexpr int ES_THREADS = 64;
constexpr int ES_SLOTS = 136;  // live faces = 2V − 4 ≤ 132
#ifndef PK_ES_KEYS
#define PK_ES_KEYS 92
#endif
#ifndef PK_ES_HORIZON
#define PK_ES_HORIZON 24
#endif
#ifndef PK_ES_HCAP
#define PK_ES_HCAP 30
#endif
constexpr int ES_KEYS = PK_ES_KEYS; // slots with a float key in shared memory (the rest is scanned from the slab);
                               // 92/4 is odd: a thread's keys are contiguous and 128-bit loads are conflict-free
constexpr int ES_VERTS = 68;   // 4 + 64 iterations
constexpr int ES_HORIZON = PK_ES_HORIZON; // longer horizons are handed back: at 16 one pair of the 1 M-body scene was, and
                                          // epa_kernel spent 0.45 ms of every step on it alone (observed max: 17)
constexpr int ES_STACK = 8;    // observed max 4
constexpr int ES_HEAP_MAX = EPA_MAX_FACES; // HEAP mode: heap entries (= faces ever created), as epa_kernel
// per-thread slab: planes, topology, vertices (+ in HEAP mode the heap entries beyond the shared-memory top)
constexpr int ES_GKEYS = (ES_SLOTS - ES_KEYS + 3) / 4 * 4; // float keys of the slots beyond shared memory
__host__ __device__ constexpr size_t es_slab_bytes(bool heap)
{
    return static_cast<size_t>(ES_SLOTS) * (32 + 8) + static_cast<size_t>(ES_VERTS) * (32 + 48) + static_cast<size_t>(ES_GKEYS) * 4 +
           (heap ? static_cast<size_t>(ES_HEAP_MAX) * (8 + 4) : 0);
}

constexpr int ES_HCAP = PK_ES_HCAP; // heap entries of the exact-heap mode kept in shared memory (polyhedron pairs: observed max 34)

// How pop_face finds the closest live face, chosen per pair.  SCAN (pairs with a sphere): float keys, no
// heap — distances of distinct faces practically never tie.  HEAP (polyhedron pairs): the reference's
// binary heap restated in shared memory — coplanar faces tie all the time, but these polytopes stay small
// enough for the whole heap to fit.  Both share one area of shared memory.
// per-thread area: ES_KEYS floats (SCAN) or ES_HCAP × (double distance, uint8 face) (HEAP).  Threads of one
// block are in different modes at the same time, so both layouts keep a thread inside its own bytes.
struct alignas(16) EsPopThread
{
    union
    {
        float key[ES_KEYS]; // float(distance) rounded down; +inf = free slot; scanned four at a time
        struct
        {
            double hd[ES_HCAP];   // heap: copy of the face distance
            uint32_t hf[ES_HCAP]; // heap: slot | creation serial << 8 (slots are recycled: the serial tells a
                                  // lazily deleted entry from the face that lives in its slot now)
        } heap;
    };
};
static_assert(sizeof(EsPopThread) == ES_KEYS * 4 && ES_KEYS % 8 == 4, "the heap must fit the key area; ES_KEYS/4 must be odd");
struct EsPop
{
    EsPopThread th[ES_THREADS];
};

struct EsSmem
{
    EsPop pop;
    double f[2][10][ES_THREADS];     // shape views: p xyz, h xyz, q xyzw
    uint32_t vert_off[2][ES_THREADS]; // HULL: first vertex in the context's vertex pool
    float hull_r[2][ES_THREADS];
    int kind[2][ES_THREADS];
    uint32_t nverts[2][ES_THREADS];
    uint32_t hz[ES_HORIZON][ES_THREADS];   // start:7 | end:7 | adjacent slot:8 | its edge:2 | new slot:8
    uint16_t ring[ES_HORIZON][ES_THREADS]; // successor slot | predecessor slot << 8 (0xFF = none)
};
static_assert(sizeof(EsSmem) <= 48 * 1024, "EsSmem must fit static shared memory");

struct EsSlab
{
    double *plane;            // normal xyz, distance: one 32-byte sector per slot
    unsigned long long *topo; // bytes 0-2 vertices, 3-5 adjacent slots (0xFF = none)
    double *vpos;             // p = pa − pb, padded to 32 bytes
    double *vab;              // pa xyz, pb xyz
    float *gkey;              // SCAN mode: float keys of slots ≥ ES_KEYS (polytopes past ≈45 iterations)
    double *hd;               // HEAP mode: heap entries ≥ ES_HCAP (distance)
    uint32_t *hf;             // HEAP mode: heap entries ≥ ES_HCAP (slot | creation serial << 8)
    __device__ __forceinline__ EsSlab(unsigned char *base, bool heap)
    {
        const size_t slots = ES_SLOTS;
        plane = reinterpret_cast<double *>(base);
        base += slots * 32;
        topo = reinterpret_cast<unsigned long long *>(base);
        base += slots * 8;
        vpos = reinterpret_cast<double *>(base);
        base += static_cast<size_t>(ES_VERTS) * 32;
        vab = reinterpret_cast<double *>(base);
        base += static_cast<size_t>(ES_VERTS) * 48;
        gkey = reinterpret_cast<float *>(base);
        base += static_cast<size_t>(ES_GKEYS) * 4;
        hd = reinterpret_cast<double *>(base);
        base += static_cast<size_t>(ES_HEAP_MAX) * 8;
        hf = reinterpret_cast<uint32_t *>(base);
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
__device__ __forceinline__ int hz_start(uint32_t h) { return static_cast<int>(h & 0x7Fu); }
__device__ __forceinline__ int hz_end(uint32_t h) { return static_cast<int>((h >> 7) & 0x7Fu); }
__device__ __forceinline__ int hz_adj(uint32_t h) { return static_cast<int>((h >> 14) & 0xFFu); }
__device__ __forceinline__ int hz_e2(uint32_t h) { return static_cast<int>((h >> 22) & 0x3u); }
__device__ __forceinline__ int hz_slot(uint32_t h) { return static_cast<int>(h >> 24); }
// initial tetrahedron: faces (0,1,2|3) (0,2,3|1) (0,3,1|2) (1,3,2|0)  (collision.cpp:361-364)
__device__ __forceinline__ int fi4(int f) { return f == 3 ? 1 : 0; }
__device__ __forceinline__ int fj4(int f) { return (0x3321 >> (4 * f)) & 0xF; }
__device__ __forceinline__ int fk4(int f) { return (0x2132 >> (4 * f)) & 0xF; }
__device__ __forceinline__ int fo4(int f) { return (0x0213 >> (4 * f)) & 0xF; }
__device__ __forceinline__ void es_prefetch(const void *p)
{
#ifdef __CUDA_ARCH__
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
__device__ __forceinline__ float es_inf() { return __int_as_float(0x7F800000); }

template <class SM> __device__ __forceinline__ void es_put_shape(SM &sm, int which, const ShapeView &v, const BodyArrays &ba)
{
    const int t = threadIdx.x;
    sm.f[which][0][t] = v.p.x; sm.f[which][1][t] = v.p.y; sm.f[which][2][t] = v.p.z;
    sm.f[which][3][t] = v.h.x; sm.f[which][4][t] = v.h.y; sm.f[which][5][t] = v.h.z;
    sm.f[which][6][t] = v.q.x; sm.f[which][7][t] = v.q.y; sm.f[which][8][t] = v.q.z; sm.f[which][9][t] = v.q.w;
    sm.vert_off[which][t] = static_cast<uint32_t>(v.vf - ba.verts_f);
    sm.hull_r[which][t] = v.hull_r;
    sm.kind[which][t] = v.kind;
    sm.nverts[which][t] = v.nverts;
}
template <class SM> __device__ __forceinline__ ShapeView es_get_shape(const SM &sm, int which, const BodyArrays &ba)
{
    const int t = threadIdx.x;
    ShapeView v;
    v.p = {sm.f[which][0][t], sm.f[which][1][t], sm.f[which][2][t]};
    v.h = {sm.f[which][3][t], sm.f[which][4][t], sm.f[which][5][t]};
    v.q = {sm.f[which][6][t], sm.f[which][7][t], sm.f[which][8][t], sm.f[which][9][t]};
    v.verts = ba.verts + 3ull * sm.vert_off[which][t];
    v.vf = ba.verts_f + sm.vert_off[which][t];
    v.hull_r = sm.hull_r[which][t];
    v.kind = sm.kind[which][t];
    v.nverts = sm.nverts[which][t];
    return v;
}

// One copy of the face-plane arithmetic (collision.cpp:273-297) instead of one per call site: with eight
// warps per SM at unrelated program counters, instruction-cache misses were 22 % of the issue stalls.
__device__ __forceinline__ double4 es_face_plane(d3 pi, d3 pj, d3 pk, bool has_opp, d3 popp, bool &flip)
{
    d3 n;
    double dist;
    flip = epa_face_plane(pi, pj, pk, has_opp, popp, n, dist);
    return make_double4(n.x, n.y, n.z, dist);
}

// collision.cpp:424-454
__device__ __noinline__ void es_write_result(const EsSlab &sl, double4 nd, unsigned long long t, ContactRec *out, uint64_t key, ContactRec *stage)
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
    if (stage) *stage = *out; // a second copy in the thread's (now dead) shared-memory area, flushed to the host by the warp
}

// HEAP mode: the first ES_HCAP heap entries live in the thread's shared-memory area (every sift starts
// there; polyhedron pairs never leave it), deeper ones in its slab.
struct EsHeapRef
{
    EsPopThread *pt;
    double *gd;
    uint32_t *gf;
    __device__ __forceinline__ double d(int k) const { return k < ES_HCAP ? pt->heap.hd[k] : gd[k]; }
    __device__ __forceinline__ uint32_t f(int k) const { return k < ES_HCAP ? pt->heap.hf[k] : gf[k]; }
    __device__ __forceinline__ void set(int k, double dist, uint32_t face) const
    {
        if (k < ES_HCAP)
        {
            pt->heap.hd[k] = dist;
            pt->heap.hf[k] = face;
        }
        else
        {
            gd[k] = dist;
            gf[k] = face;
        }
    }
};
// libstdc++ std::__push_heap with comp(a,b) = dist[a] > dist[b]  (collision.cpp:390-395)
__device__ __forceinline__ void es_sift_up(const EsHeapRef &h, int hole, double vd, uint32_t vf)
{
    int parent = (hole - 1) / 2;
    while (hole > 0)
    {
        const double pd = h.d(parent);
        if (!(pd > vd)) break;
        h.set(hole, pd, h.f(parent));
        hole = parent;
        parent = (hole - 1) / 2;
    }
    h.set(hole, vd, vf);
}
// libstdc++ std::pop_heap (→ __pop_heap → __adjust_heap) followed by back()/pop_back()
__device__ __forceinline__ uint32_t es_heap_pop(const EsHeapRef &h, int &size)
{
    const uint32_t top = h.f(0);
    if (size == 1)
    {
        size = 0;
        return top;
    }
    const int len = size - 1;
    const double vd = h.d(len);
    const uint32_t vf = h.f(len);
    int hole = 0, child = 0;
    while (child < (len - 1) / 2)
    {
        child = 2 * (child + 1);
        const double rd = h.d(child), ld = h.d(child - 1);
        if (rd > ld)
        {
            child--;
            h.set(hole, ld, h.f(child));
        }
        else
            h.set(hole, rd, h.f(child));
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2)
    {
        child = 2 * (child + 1);
        h.set(hole, h.d(child - 1), h.f(child - 1));
        hole = child - 1;
    }
    es_sift_up(h, hole, vd, vf);
    size = len;
    return top;
}

// The initial tetrahedron of every GJK hit (build_initial_tetrahedron, collision.cpp:355-388), computed by a
// kernel of its own: four face planes and the brute-force adjacency are ≈600 instructions that all threads
// execute in step here, instead of a divergent set-up path that stalls a warp of epa_scan_kernel every time
// one of its lanes starts a pair.
struct alignas(16) EpaInit
{
    double plane[4][4];         // normal xyz, distance of faces (0,1,2|3) (0,2,3|1) (0,3,1|2) (1,3,2|0)
    unsigned long long topo[4]; // vertices (flipped where the normal faced the opposite vertex), neighbours, serial
    uint32_t bad;               // a distance is NaN / inf: not for the heap-free path
    uint32_t _pad[3];
};
static_assert(sizeof(EpaInit) == 176, "EpaInit layout");

__global__ void __launch_bounds__(128)
epa_init_kernel(const SimplexRec *__restrict__ simplices, const unsigned long long *__restrict__ hit_count_ptr, uint64_t hit_capacity,
                EpaInit *__restrict__ init)
{
    const uint64_t s = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    unsigned long long nhits = *hit_count_ptr;
    if (nhits > hit_capacity) nhits = hit_capacity;
    if (s >= nhits) return;
    const SimplexRec *r = simplices + s;
    if ((r->n & 0xFFu) != 4u) return; // padded simplices are set up by epa_kernel
    d3 pv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        SupportPt sp;
        sp.pa = d3{r->v[i][0], r->v[i][1], r->v[i][2]};
        sp.pb = d3{r->v[i][3], r->v[i][4], r->v[i][5]};
        pv[i] = P(sp);
    }
    EpaInit o;
    uint32_t tv[4]; // vertex triples, one byte each
    bool bad = false;
#pragma unroll
    for (int f = 0; f < 4; ++f)
    {
        bool flip;
        const double4 pl = es_face_plane(pv[fi4(f)], pv[fj4(f)], pv[fk4(f)], true, pv[fo4(f)], flip);
        tv[f] = static_cast<uint32_t>(fi4(f)) | (static_cast<uint32_t>(flip ? fk4(f) : fj4(f)) << 8) |
                (static_cast<uint32_t>(flip ? fj4(f) : fk4(f)) << 16);
        o.plane[f][0] = pl.x;
        o.plane[f][1] = pl.y;
        o.plane[f][2] = pl.z;
        o.plane[f][3] = pl.w;
        if (!(fabs(pl.w) < 1e30)) bad = true; // NaN / inf: the key order would not be the heap's
    }
    // brute-force adjacency (collision.cpp:373-388).  An undirected tetrahedron edge belongs to exactly two
    // faces, so a directed edge has at most one reversed partner and the reference's i<j visiting order
    // cannot matter: each face looks its three partners up independently.
#pragma unroll
    for (int f = 0; f < 4; ++f)
    {
        unsigned long long w = tv[f];
#pragma unroll
        for (int e1 = 0; e1 < 3; ++e1)
        {
            const uint32_t u1 = (tv[f] >> (8 * e1)) & 0xFFu, v1 = (tv[f] >> (8 * ((e1 + 1) % 3))) & 0xFFu;
            unsigned long long adj = 0xFFull;
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const uint32_t q = tv[j];
                const bool m = (j != f) && (((q & 0xFFu) == v1 && ((q >> 8) & 0xFFu) == u1) ||
                                            (((q >> 8) & 0xFFu) == v1 && ((q >> 16) & 0xFFu) == u1) ||
                                            (((q >> 16) & 0xFFu) == v1 && (q & 0xFFu) == u1));
                if (m) adj = static_cast<unsigned long long>(j);
            }
            w |= adj << (24 + 8 * e1);
        }
        o.topo[f] = w | (static_cast<unsigned long long>(f) << 48); // creation serial
    }
    o.bad = bad ? 1u : 0u;
    o._pad[0] = o._pad[1] = o._pad[2] = 0;
    init[s] = o;
}

// 5 blocks of 64 threads per SM (≤ 204 registers, 5 × 42 KB of shared memory): 15.6 → 15.1 ms at 1 M bodies against 4
#ifndef PK_ES_MIN_BLOCKS
#define PK_ES_MIN_BLOCKS 5
#endif
#ifndef PK_ES_FETCH_MIN
#define PK_ES_FETCH_MIN 12
#endif

// The hit list is grouped by cost class (order[]: sphere–sphere, sphere–polyhedron, polyhedron–polyhedron);
// class_count[] = sizes of the groups with 0, 1, 2 spheres.  The SCAN instance takes the two groups with a
// sphere in one persistent launch, the HEAP instance the polyhedron pairs (a single kernel choosing the
// mode per pair was measured slower: more registers, and both pop paths in every warp's instruction
// stream).  Pairs handed back are published in fallback_list[] (entries start as EPA_LIST_EMPTY), where a
// small epa_kernel launch running next to these kernels picks them up (pk_api.cu).
// MIRROR: the pk_collide instance, which also delivers finished records to the caller's pinned buffer.
template <bool HEAP, bool MIRROR>
__global__ void __launch_bounds__(ES_THREADS, PK_ES_MIN_BLOCKS)
epa_scan_kernel(BodyArrays bodies, const uint64_t *__restrict__ keys, const uint32_t *__restrict__ pair_a,
                const uint32_t *__restrict__ pair_b, const SimplexRec *__restrict__ simplices,
                const unsigned long long *__restrict__ hit_count_ptr, uint64_t hit_capacity,
                const uint32_t *__restrict__ out_index, const uint32_t *__restrict__ order, ContactRec *__restrict__ contacts,
                uint8_t *__restrict__ valid, unsigned char *__restrict__ slabs, unsigned long long *__restrict__ cursor,
                unsigned long long *__restrict__ counters /* [0]=valid contacts */, uint32_t *__restrict__ fallback_list,
                unsigned long long *__restrict__ fallback_count, const unsigned long long *__restrict__ class_count,
                const uint32_t *__restrict__ leftovers, const unsigned long long *__restrict__ leftover_count,
                const EpaInit *__restrict__ init, ContactRec *contacts_host)
{
    // Work of the HEAP instance: first the pairs the SCAN instance handed back (leftovers[], complete at
    // launch: mostly sphere–sphere pairs with an exact distance tie, long ones — started first so that
    // they overlap the rest), then the polyhedron pairs of order[].
    __shared__ EsSmem shm;
    const int t = threadIdx.x;
    const uint64_t tid = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    const EsSlab sl(slabs + tid * es_slab_bytes(HEAP), HEAP);
    const EsHeapRef hp{&shm.pop.th[t], sl.hd, sl.hf};
    unsigned long long nhits = *hit_count_ptr;
    if (nhits > hit_capacity) nhits = hit_capacity;
    const unsigned long long fb_capacity = nhits;
    unsigned long long first_hit = class_count[2] + class_count[1]; // where the polyhedron pairs start
    if (first_hit > nhits) first_hit = nhits;
    unsigned long long nleft = 0;
    if (HEAP)
    {
        nhits -= first_hit;
        nleft = *leftover_count < fb_capacity ? *leftover_count : fb_capacity;
    }
    else
    {
        nhits = first_hit;
        first_hit = 0;
    }
    const float INF = es_inf();

    bool active = false, done = false;
    int nverts = 0, iter = 0, hi = 0; // hi: slots [0, hi) have been used by the current polytope
    bool keys_dirty = true;           // the key area does not hold +inf beyond the current polytope
    int gdirty = ES_SLOTS;            // slab keys [ES_KEYS, gdirty) may hold something else than +inf
    unsigned long long fm0 = 0, fm1 = 0, fm2 = 0; // free slots, 64 per word
    uint32_t out_slot = 0, cur_sidx = 0;
    uint64_t key = 0;
    // tie breaking (see pop): lower bound of the distances of the lazily deleted heap entries the reference's
    // heap would still hold (faces killed by the flood fill; a pop removes every entry below the popped one)
    double stale_lb = 1e300;
    int batch_n = 0; // hz[0, batch_n) still describes the last batch of faces, in push order
    unsigned long long n_valid = 0;
    int heap_size = 0, nfaces = 0; // HEAP: number of faces created so far = serial of the next one
    constexpr unsigned FULL = 0xFFFFFFFFu;

    int fb = 0; // reason + 1 when the current pair has to go to epa_kernel (one atomic site for all of them)
    bool flush_pending = false; // a finished record waits in this lane's shared-memory area (pk_collide)
    uint32_t flush_slot = 0;
    auto is_free = [&](int f) -> bool
    {
        const unsigned long long w = (f < 64) ? fm0 : (f < 128 ? fm1 : fm2);
        return (w >> (f & 63)) & 1ull;
    };
    auto kill_slot = [&](int f)
    {
        if constexpr (!HEAP)
        {
            if (f < ES_KEYS)
                shm.pop.th[t].key[f] = INF;
            else
                sl.gkey[f - ES_KEYS] = INF;
        }
        if (f < 64)
            fm0 |= 1ull << f;
        else if (f < 128)
            fm1 |= 1ull << (f - 64);
        else
            fm2 |= 1ull << (f - 128);
    };
    // float key of slot s: shared memory for the first ES_KEYS slots, derived from the slab beyond
    auto key_of = [&](int s) -> float { return (s < ES_KEYS) ? shm.pop.th[t].key[s] : sl.gkey[s - ES_KEYS]; };
    // slot of a new face.  SCAN: lowest free first (keeps the live polytope dense at the start of the slab
    // and inside the keyed slots).  HEAP: never-used slots first, so that as long as the polytope has
    // created at most ES_SLOTS faces a heap entry's slot equals its serial and "slot free" is the whole
    // obsolete test; only bigger polytopes recycle and have to compare serials.
    auto take_slot = [&]() -> int
    {
        int slot;
        if (HEAP && nfaces < ES_SLOTS)
            slot = nfaces;
        else if (fm0)
            slot = __ffsll(static_cast<long long>(fm0)) - 1;
        else if (fm1)
            slot = 64 + __ffsll(static_cast<long long>(fm1)) - 1;
        else
            slot = 128 + __ffsll(static_cast<long long>(fm2)) - 1;
        if (slot < 64)
            fm0 &= ~(1ull << slot);
        else if (slot < 128)
            fm1 &= ~(1ull << (slot - 64));
        else
            fm2 &= ~(1ull << (slot - 128));
        return slot;
    };

    for (;;)
    {
        if (fb)
        {
            unsigned long long i = atomicAdd(fallback_count, 1ull);
            if (i < fb_capacity) fallback_list[i] = cur_sidx; // one word: publication needs no fence
#ifdef PK_ES_REASONS
            atomicAdd(fallback_count - (HEAP ? 1 : 0) + 9 + fb, 1ull); // debug build only: C_EPA_REASONS + reason
            if (!HEAP) atomicAdd(fallback_count + 14, static_cast<unsigned long long>(iter));
#endif
            fb = 0;
            active = false;
        }
        const unsigned m_active = __ballot_sync(FULL, active);
        const unsigned m_idle = __ballot_sync(FULL, !active && !done);
        if constexpr (MIRROR)
        {
            // pk_collide: finished records go to the caller's pinned buffer from here, so that no device→host
            // copy of the contacts has to wait for the kernels to end.  The lane that finished a pair left the
            // record in its shared-memory area; eleven lanes store it with ONE 8-byte store each (88
            // contiguous bytes).  Letting the finishing lane store to host memory itself — eleven dependent
            // stores from one lane of a divergent warp — slowed the kernel by more than the copy costs.
            unsigned m_flush = __ballot_sync(FULL, flush_pending);
            if (m_flush)
            {
                __syncwarp();
                const int lane = t & 31;
                while (m_flush)
                {
                    const int L = __ffs(static_cast<int>(m_flush)) - 1;
                    m_flush &= m_flush - 1u;
                    const uint32_t slot_l = __shfl_sync(FULL, flush_slot, L);
                    if (lane < 11)
                        reinterpret_cast<unsigned long long *>(contacts_host + slot_l)[lane] =
                            reinterpret_cast<const unsigned long long *>(&shm.pop.th[(t & ~31) + L])[lane];
                }
                __syncwarp();
                flush_pending = false;
            }
        }
        if (m_active == 0 && m_idle == 0) break;
        if (!active && !done && (__popc(m_idle) >= PK_ES_FETCH_MIN || m_active == 0))
        {
            unsigned long long slot = atomicAdd(cursor, 1ull);
            if (slot >= nleft + nhits)
                done = true;
            else
            {
                cur_sidx = (slot < nleft) ? leftovers[slot] : order[first_hit + (slot - nleft)];
                const SimplexRec *r = simplices + cur_sidx;
                const uint32_t pair = r->pair;
                iter = 0;
                if ((r->n & 0xFFu) != 4u)
                    fb = 1; // pad_simplex path (collision.cpp:191-248): rare, left to epa_kernel
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
                        es_put_shape(shm, 0, A, bodies);
                        es_put_shape(shm, 1, B, bodies);
                    }
                    const EpaInit *in = init + cur_sidx; // epa_init_kernel: planes and topology of the tetrahedron
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                    {
                        SupportPt s;
                        s.pa = d3{r->v[i][0], r->v[i][1], r->v[i][2]};
                        s.pb = d3{r->v[i][3], r->v[i][4], r->v[i][5]};
                        sl.set_vert(i, s, P(s));
                    }
                    if constexpr (!HEAP)
                    {
                        // leftovers of the previous polytope; everything at start and after a HEAP pair,
                        // whose heap occupied the same shared memory
                        const int dirty = keys_dirty ? ES_KEYS : (hi < ES_KEYS ? hi : ES_KEYS);
                        for (int s = 4; s < dirty; ++s) shm.pop.th[t].key[s] = INF;
                        for (int s = ES_KEYS; s < gdirty; ++s) sl.gkey[s - ES_KEYS] = INF;
                        gdirty = ES_KEYS;
                        keys_dirty = false;
                    }
                    else
                        keys_dirty = true;
                    heap_size = 0;
                    const bool bad = in->bad != 0u;
#pragma unroll
                    for (int f = 0; f < 4; ++f)
                    {
                        const double2 *q = reinterpret_cast<const double2 *>(in->plane[f]);
                        const double2 n01 = q[0], n2d = q[1];
                        sl.store_plane(f, d3{n01.x, n01.y, n2d.x}, n2d.y);
                        sl.topo[f] = in->topo[f];
                        if constexpr (HEAP)
                        {
                            es_sift_up(hp, heap_size, n2d.y, static_cast<uint32_t>(f) | (static_cast<uint32_t>(f) << 8)); // push_face
                            ++heap_size;
                        }
                        else
                        {
                            shm.pop.th[t].key[f] = __double2float_rd(n2d.y);
                            shm.hz[f][t] = static_cast<uint32_t>(f) << 24; // push order of this batch (tie breaking)
                        }
                    }
                    fm0 = ~0xFull;
                    fm1 = ~0ull;
                    fm2 = (1ull << (ES_SLOTS - 128)) - 1ull;
                    hi = 4;
                    nfaces = 4;
                    nverts = 4;
                    stale_lb = 1e300;
                    batch_n = 4;
                    active = true;
                    if (bad) fb = 4;
                }
            }
        }
        if (!active || fb) continue;

        // ---- pop_face (collision.cpp:397-408) without a heap: the live face with the smallest distance ----
        int min_face = -1;
        if constexpr (HEAP)
        {
            while (heap_size > 0) // skip obsolete entries: slot free, or re-used by a younger face
            {
                const uint32_t id = es_heap_pop(hp, heap_size);
                const int f = static_cast<int>(id & 0xFFu);
                if (is_free(f)) continue;
                if (nfaces > ES_SLOTS && static_cast<uint32_t>(sl.topo[f] >> 48) != (id >> 8)) continue;
                min_face = f;
                break;
            }
        }
        else
        {
            float m = INF;
            int cnt = 0, second = -1;
            const int hs = hi < ES_KEYS ? hi : ES_KEYS;
            const float4 *kq = reinterpret_cast<const float4 *>(shm.pop.th[t].key);
            for (int s = 0; s < hs; s += 4) // slots in [hs, s+4) hold +inf: they never win and never count
            {
                const float4 k4 = kq[s >> 2];
                const float kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                for (int u = 0; u < 4; ++u)
                {
                    if (kk[u] < m)
                    {
                        m = kk[u];
                        min_face = s + u;
                        cnt = 1;
                    }
                    else if (kk[u] == m)
                    {
                        if (cnt == 1) second = s + u;
                        ++cnt;
                    }
                }
            }
            if (hi > ES_KEYS) // only polytopes past ≈45 iterations: same scan over the keys kept in the slab
            {
                const float4 *gq = reinterpret_cast<const float4 *>(sl.gkey);
                for (int s = ES_KEYS; s < hi; s += 4)
                {
                    const float4 k4 = gq[(s - ES_KEYS) >> 2];
                    const float kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                    {
                        if (kk[u] < m)
                        {
                            m = kk[u];
                            min_face = s + u;
                            cnt = 1;
                        }
                        else if (kk[u] == m)
                        {
                            if (cnt == 1) second = s + u;
                            ++cnt;
                        }
                    }
                }
            }
            if (min_face >= 0 && cnt > 1)
            {
                // several live faces share the minimal float key (mirror-image faces of sphere–sphere polytopes
                // differ in the last bits only): compare the exact distances
                double best = sl.plane[4 * min_face + 3];
                bool tie = false;
                if (cnt == 2)
                {
                    const double d = sl.plane[4 * second + 3];
                    if (d < best)
                    {
                        best = d;
                        min_face = second;
                    }
                    else if (d == best)
                        tie = true;
                }
                else
                {
                    for (int s = min_face + 1; s < hi; ++s)
                    {
                        if (key_of(s) != m) continue;
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
                }
                if (tie)
                {
                    // Which of several equidistant faces std::pop_heap delivers depends on the heap's history.
                    // Two cases can be decided without it.  Let d be the tied minimum and suppose every
                    // lazily deleted entry still in the heap is farther than d (stale_lb, below).
                    //  (1) exactly one tied face is older than the last batch of pushes: before that batch it
                    //      was the strict minimum of the whole heap, hence the root; __push_heap moves a parent
                    //      down only if parent > value, so no new entry with the same distance passed it.
                    //  (2) all tied faces belong to the last batch: every older entry is farther than d, so the
                    //      first of them to be pushed sifted up to the root and, as in (1), stayed there.
                    // (2) covers the degenerate faces (normal 0, distance 0, collision.cpp:282-287) that one
                    // iteration creates in pairs; (1) the mirror-image faces of sphere–sphere polytopes.
                    bool resolved = false;
                    if (best < stale_lb)
                    {
                        int n_old = 0, old_face = -1, first_new = -1;
                        for (int e = batch_n - 1; e >= 0; --e)
                        {
                            const int s = hz_slot(shm.hz[e][t]);
                            if (!is_free(s) && sl.plane[4 * s + 3] == best) first_new = s;
                        }
                        for (int s = 0; s < hi; ++s)
                        {
                            if (key_of(s) != m || sl.plane[4 * s + 3] != best) continue; // (free slots have key +inf)
                            bool in_batch = false;
                            for (int e = 0; e < batch_n; ++e) in_batch = in_batch || hz_slot(shm.hz[e][t]) == s;
                            if (!in_batch)
                            {
                                ++n_old;
                                old_face = s;
                            }
                        }
                        if (n_old == 1)
                        {
                            min_face = old_face;
                            resolved = true;
                        }
                        else if (n_old == 0 && first_new >= 0)
                        {
                            min_face = first_new;
                            resolved = true;
                        }
                    }
                    if (!resolved)
                    {
                        fb = 2;
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
        if (mf.w > stale_lb) stale_lb = mf.w; // entries closer than the popped face have left the heap
        bool finished = iter >= 64; // best guess after the loop (collision.cpp:500-503)
        if (!finished) ++iter;
        // the flood fill starts with the three neighbours of this face: have them on their way during the
        // support evaluation
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            const int b = es_adj(mt, k);
            if (b != 0xFF)
            {
                es_prefetch(sl.plane + 4 * b);
                es_prefetch(sl.topo + b);
            }
        }
        const d3 mn{mf.x, mf.y, mf.z};
        SupportPt sp{}; // minkowski_support (collision.h:41-49); one copy of the support code for both shapes
        if (!finished)
        {
            {
                const ShapeView A = es_get_shape(shm, 0, bodies);
                sp.pa = support(A, mn);
            }
            {
                const ShapeView B = es_get_shape(shm, 1, bodies);
                sp.pb = support(B, -mn);
            }
        }
        const d3 p = P(sp);
        if (finished || dot(mn, p) - mf.w < 1e-6) // converged (collision.cpp:465-466)
        {
            es_write_result(sl, mf, mt, contacts + out_slot, key, MIRROR ? reinterpret_cast<ContactRec *>(&shm.pop.th[t]) : nullptr);
            if constexpr (MIRROR)
            {
                flush_pending = true;
                flush_slot = out_slot;
                if (hi < 22) hi = 22; // the staged record covers the first 22 key slots: the next fetch must reset them
            }
            valid[out_slot] = 1;
            ++n_valid;
            active = false;
            continue;
        }

        // ---- find_silhouette (collision.cpp:315-353): LIFO flood fill, edge order preserved ----
        bool bad = false;
#ifdef PK_ES_WHY
        unsigned why = 0;
#endif
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
                    live[i] = a != 0xFF && !is_free(a);
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
                    if (is_free(a)) continue; // reached through two edges of this face: the first visit killed it
                    if (dot(d3{nf[i].x, nf[i].y, nf[i].z}, p) > nf[i].w + 1e-6)
                    {
                        kill_slot(a);
                        if (nf[i].w < stale_lb) stale_lb = nf[i].w; // its heap entry stays behind
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
                        {
                            bad = true;
                            PK_ES_WHY_TAG(1);
                        }
                    }
                    else
                    {
                        // horizon edge (cur.v[i], cur.v[i+1], a).  link_faces(f, a, start, end) will need the
                        // edge of `a` that starts at `end` (collision.cpp:305-313): `a` is in registers now
                        const int st = es_v(cur, i), en = es_v(cur, (i + 1) % 3);
                        const int e2 = (es_v(nt[i], 0) == en) ? 0 : (es_v(nt[i], 1) == en ? 1 : 2);
                        if (es_v(nt[i], e2) != en)
                        {
                            bad = true; // unmatched link: slot recycling is no longer safe
                            PK_ES_WHY_TAG(2);
                        }
                        if (nh < ES_HORIZON)
                        {
                            shm.hz[nh][t] = static_cast<uint32_t>(st) | (static_cast<uint32_t>(en) << 7) | (static_cast<uint32_t>(a) << 14) |
                                            (static_cast<uint32_t>(e2) << 22);
                            ++nh;
                        }
                        else
                        {
                            bad = true;
                            PK_ES_WHY_TAG(4);
                        }
                        es_prefetch(sl.vpos + 4 * st); // operands of the face loop
                        es_prefetch(sl.vpos + 4 * en);
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
        const int nfree = __popcll(fm0) + __popcll(fm1) + __popcll(fm2);
        const bool full = nfree < nh || nverts >= ES_VERTS || (HEAP && heap_size + nh > ES_HEAP_MAX);
        if (bad || nh < 3 || full)
        {
#ifdef PK_ES_WHY
            if (HEAP) printf("[why] early bad=%d why=%u nh=%d full=%d iter=%d nverts=%d\n", (int)bad, why, nh, (int)full, iter, nverts);
#endif
            fb = full ? 3 : 4;
            continue;
        }
        sl.set_vert(nverts, sp, p);
        const int p_idx = nverts++;
        // new faces (start, end, p_idx), no orientation flip (collision.cpp:475-482), slots lowest free first;
        // the vertex loads of edge e+1 are issued before the arithmetic of edge e
        unsigned long long end_seen0 = 0, start_seen0 = 0;
        uint32_t end_seen1 = 0, start_seen1 = 0;
        uint32_t h = shm.hz[0][t];
        d3 ps = sl.vp(hz_start(h)), pe = sl.vp(hz_end(h));
        for (int e = 0; e < nh; ++e)
        {
            const uint32_t hc = h;
            const d3 cs = ps, ce = pe;
            if (e + 1 < nh)
            {
                h = shm.hz[e + 1][t];
                ps = sl.vp(hz_start(h));
                pe = sl.vp(hz_end(h));
            }
            const int st = hz_start(hc), en = hz_end(hc);
            const int slot = take_slot();
            if (slot >= hi) hi = slot + 1;
            bool flip_unused;
            const double4 pl = es_face_plane(cs, ce, p, false, d3{0, 0, 0}, flip_unused);
            const d3 n{pl.x, pl.y, pl.z};
            const double dist = pl.w;
            sl.store_plane(slot, n, dist);
            if (!(fabs(dist) < 1e30))
            {
                bad = true;
                PK_ES_WHY_TAG(8);
            }
            if constexpr (HEAP)
            {
                es_sift_up(hp, heap_size, dist, static_cast<uint32_t>(slot) | (static_cast<uint32_t>(nfaces) << 8)); // push_face, horizon order
                ++heap_size;
            }
            else
            {
                if (slot < ES_KEYS)
                    shm.pop.th[t].key[slot] = __double2float_rd(dist);
                else
                {
                    sl.gkey[slot - ES_KEYS] = __double2float_rd(dist);
                    if (slot >= gdirty) gdirty = slot + 1;
                }
            }
            ++nfaces;
            shm.hz[e][t] = hc | (static_cast<uint32_t>(slot) << 24);
            shm.ring[e][t] = 0xFFFFu;
            sl.set_adj(hz_adj(hc), hz_e2(hc), slot); // link_faces(f, adj_face, start, end), the old face's side
            // proper horizon: every vertex starts at most one edge and ends at most one, no self loop
            const bool ends_twice = (en < 64) ? ((end_seen0 >> en) & 1ull) : ((end_seen1 >> (en - 64)) & 1u);
            const bool starts_twice = (st < 64) ? ((start_seen0 >> st) & 1ull) : ((start_seen1 >> (st - 64)) & 1u);
            if (st == en || starts_twice || ends_twice)
            {
                bad = true;
                PK_ES_WHY_TAG(16);
            }
            if (en < 64)
                end_seen0 |= 1ull << en;
            else
                end_seen1 |= 1u << (en - 64);
            if (st < 64)
                start_seen0 |= 1ull << st;
            else
                start_seen1 |= 1u << (st - 64);
        }
        // ring links among the new faces (collision.cpp:484-497): face e = (start, end, p_idx) gets its
        // successor (the edge starting at `end`) on edge 1 and is that face's neighbour on edge 2
        // (starts are unique, so the successor of an edge is found by a search over the ≤16 horizon records)
        for (int e = 0; e < nh; ++e)
        {
            const uint32_t he = shm.hz[e][t];
            const int en = hz_end(he);
            const bool has_succ = (en < 64) ? ((start_seen0 >> en) & 1ull) : ((start_seen1 >> (en - 64)) & 1u);
            if (!has_succ) continue;
            int j = 0;
            uint32_t hj = shm.hz[0][t];
            while (hz_start(hj) != en) hj = shm.hz[++j][t];
            if (hz_end(hj) == hz_start(he))
            {
                bad = true; // 2-cycle: the reference links it one way only
                PK_ES_WHY_TAG(32);
            }
            shm.ring[e][t] = static_cast<uint16_t>((shm.ring[e][t] & 0xFF00u) | (hj >> 24));
            shm.ring[j][t] = static_cast<uint16_t>((shm.ring[j][t] & 0x00FFu) | ((he >> 24) << 8));
        }
        for (int e = 0; e < nh; ++e)
        {
            const uint32_t he = shm.hz[e][t];
            const uint32_t rg = shm.ring[e][t];
            sl.topo[hz_slot(he)] = static_cast<unsigned long long>(hz_start(he)) | (static_cast<unsigned long long>(hz_end(he)) << 8) |
                                   (static_cast<unsigned long long>(p_idx) << 16) | (static_cast<unsigned long long>(hz_adj(he)) << 24) |
                                   (static_cast<unsigned long long>(rg & 0xFFu) << 32) | (static_cast<unsigned long long>(rg >> 8) << 40) |
                                   (static_cast<unsigned long long>(nfaces - nh + e) << 48); // creation serial
        }
        if (bad)
        {
#ifdef PK_ES_WHY
            if (HEAP) printf("[why] late why=%u nh=%d iter=%d nverts=%d nfaces=%d\n", why, nh, iter, nverts, nfaces);
#endif
            fb = 4;
            continue;
        }
        batch_n = nh;
    }
    if (n_valid) atomicAdd(counters + 0, n_valid);
#ifdef PK_ES_WHY
    if (threadIdx.x == 0) // when does each block retire: how long is the tail of the persistent launch
    {
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        printf("[exit] %d %d %llu\n", HEAP ? 1 : 0, static_cast<int>(blockIdx.x), now);
    }
#endif
}

} // namespace pk
