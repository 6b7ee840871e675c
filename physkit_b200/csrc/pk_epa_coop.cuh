// pk_epa_coop.cuh — EPA (reference src/collision.cpp:251-509): one lane per pair for the sequential part of an
// iteration, the whole warp for the part that is parallel over horizon edges, one persistent launch for all pairs.
//
// An EPA iteration splits into
//   S1 (per pair, sequential): pop the closest face, support point, convergence test, flood fill → horizon
//      (collision.cpp:397-408, 461-470, 315-353), slot assignment for the new faces;
//   S2 (per horizon edge, independent): face plane of (start, end, p), key, link to the face across the horizon,
//      ring links to the two neighbouring new faces (collision.cpp:273-297, 305-313, 475-497).
// S1 runs one lane per pair.  S2 is dealt out evenly: the edges of all pairs of the warp are numbered by a prefix
// sum and lane l takes edges l, l+32, …, whoever owns them — lanes whose pair has finished, whose horizon is short
// or which have no pair at all (the tail of the launch) work for the others, and no lane executes anything
// redundantly.  The owner publishes its horizon, the new vertex and the slots in shared memory; the polytope stays
// in the owner's slab in HBM / L2, which any lane of the warp can address.  (ncu, 1 M bodies: the face loop ran at
// 13 of 32 lanes when every lane did its own edges, S2 runs at 25.6; profiles/r2_epa_coop_sections.md.)
//
// Two ways to pop the closest face, chosen per pair class (the hit list is grouped by class, order[]):
//   SCAN (pairs with a sphere or a many-vertex hull): no heap.  The reference pops the live face with the smallest
//      distance; which face that is depends on the heap's internal order only when two live faces tie for the
//      minimum.  A float key (distance rounded down) per face slot lives in shared memory; pop is a scan for the
//      minimal key (per 16-byte chunk: minimum of four, compared with the running minimum), exact distances are
//      consulted only when several slots share the minimal key, and an exact tie is decided where it provably can be
//      (see the proof at the tie code) or the pair is handed to the HEAP mode.
//   HEAP (polyhedron pairs, where coplanar faces tie all the time, + what SCAN handed back): libstdc++'s
//      __push_heap / __adjust_heap restated, the first ES_HCAP entries in the same shared-memory area.
// The mode is a property of the lane's current pair: all hits are in ONE list (sphere–sphere first, polyhedron pairs
// last) behind one cursor, one persistent launch.  A SCAN pair that meets a tie only the heap can break is started
// again from its simplex in HEAP mode by the same lane, at once.  (Tried and measured slower, profiles/r2_epa_*:
// one launch per mode — each ends in a tail of its own, 1.3 + 1.2 ms at 1 M bodies, and the handed-back pairs, long
// ones, only start when the first launch is over; one launch with a SCAN pass and a HEAP pass per warp — two
// copies of the loop compete for the instruction cache and the hand-backs still start late.)  The short polyhedron
// pairs at the end of the list fill the lanes while the last long sphere pairs finish.  What cannot be done here
// (padded simplices, improper horizons, horizons beyond ES_HORIZON edges, polytopes past the slab) goes to
// epa_kernel (pk_narrowphase.cuh) through a hand-back list.
//
// Face slots are recycled (a face made obsolete frees its slot, new faces take the lowest free one), so the live
// polytope (2V−4 faces) stays dense at the start of a small slab; recycling is only safe on a proper manifold:
// an unmatched link or an improper horizon hands the pair back.  Finished pairs are parked and their records
// written when the warp refills, so that the result path runs for several lanes at once.
// Results are bit-identical to epa_kernel and to the oracle.
#pragma once

#include "pk_narrowphase.cuh"
namespace pk
{

#ifdef PK_EC_TIMING // debug build: globaltimer at entry and exit of every warp of the last launch (pk_destroy prints them)
__device__ unsigned long long g_ec_times[8192][2];
#endif

constexpr int ES_THREADS = 64;
constexpr int ES_SLOTS = 136; // live faces = 2V − 4 ≤ 132
#ifndef PK_ES_KEYS
#define PK_ES_KEYS 92
#endif
#ifndef PK_ES_HORIZON
#define PK_ES_HORIZON 24
#endif
#ifndef PK_ES_HCAP
#define PK_ES_HCAP 22
#endif
constexpr int ES_KEYS = PK_ES_KEYS; // slots with a float key in shared memory (the rest is scanned from the slab);
                                    // 92/4 is odd: a thread's keys are contiguous and 128-bit loads are conflict-free
constexpr int ES_VERTS = 68;        // 4 + 64 iterations
constexpr int ES_HORIZON = PK_ES_HORIZON; // longer horizons are handed back (observed max at 1 M bodies: 17)
constexpr int ES_STACK = 8;               // flood-fill stack (observed max 4)
constexpr int ES_HEAP_MAX = EPA_MAX_FACES; // HEAP mode: heap entries (= faces ever created), as epa_kernel
constexpr int ES_HCAP = PK_ES_HCAP;        // HEAP mode: heap entries kept in shared memory (deeper ones in the slab)
constexpr int ES_DNEW = (ES_KEYS * 4 - ES_HCAP * 12) / 8; // HEAP mode: distances of the new faces, S2 → owner's pushes
constexpr int ES_GKEYS = (ES_SLOTS - ES_KEYS + 3) / 4 * 4; // float keys of the slots beyond shared memory
// per-thread slab: planes, topology, vertices, float keys beyond shared memory, heap entries beyond shared memory
// A whole number of 128-byte lines: a slab shares no L2 line with its neighbours (20 272 → 20 352 bytes: EPA 12.8 → 12.55 ms,
// DRAM reads 4.9 → 3.8 GB, writes 5.6 → 5.1 GB per launch).
__host__ __device__ constexpr size_t es_slab_bytes()
{
    return (static_cast<size_t>(ES_SLOTS) * (32 + 8) + static_cast<size_t>(ES_VERTS) * (32 + 48) + static_cast<size_t>(ES_GKEYS) * 4 +
            static_cast<size_t>(ES_HEAP_MAX) * (8 + 4) + 127) / 128 * 128;
}

// per-thread pop area: ES_KEYS floats (SCAN) or the top of the heap + the new faces' distances (HEAP)
struct alignas(16) EsPopThread
{
    union
    {
        float key[ES_KEYS]; // float(distance) rounded down; +inf = free slot; scanned four at a time
        struct
        {
            double hd[ES_HCAP];   // heap: copy of the face distance
            double dnew[ES_DNEW]; // distances of this iteration's new faces, in horizon order
            uint32_t hf[ES_HCAP]; // heap: slot | creation serial << 8 (slots are recycled: the serial tells a
                                  // lazily deleted entry from the face that lives in its slot now)
        } heap;
    };
};
static_assert(sizeof(EsPopThread) == ES_KEYS * 4 && ES_KEYS % 8 == 4 && ES_DNEW >= 8, "pop area layout");

struct EcSmem
{
    EsPopThread pop[ES_THREADS];
    double f[2][10][ES_THREADS];      // shape views: p xyz, h xyz, q xyzw
    uint32_t vert_off[2][ES_THREADS]; // HULL: first vertex in the context's vertex pool
    float hull_r[2][ES_THREADS];
    int kind[2][ES_THREADS];
    uint32_t nverts[2][ES_THREADS];
    uint32_t hz[ES_HORIZON][ES_THREADS]; // start:7 | end:7 | adjacent slot:8 | new slot:8 (bits 24-31)
    double pnew[3][ES_THREADS];          // the support point of this iteration (apex of the new faces)
    uint32_t meta[ES_THREADS];           // index of that vertex:8 | creation serial of the first new face << 8
    uint32_t bad[ES_THREADS];            // set by any lane that finds the owner's horizon improper
};
static_assert(sizeof(EcSmem) <= 48 * 1024, "EcSmem must fit static shared memory");

struct EsSlab
{
    double *plane;            // normal xyz, distance: one 32-byte sector per slot
    unsigned long long *topo; // bytes 0-2 vertices, 3-5 adjacent slots (0xFF = none), bits 48-63 creation serial
    double *vpos;             // p = pa − pb, padded to 32 bytes
    double *vab;              // pa xyz, pb xyz
    float *gkey;              // SCAN mode: float keys of slots ≥ ES_KEYS (polytopes past ≈45 iterations)
    double *hd;               // HEAP mode: heap entries ≥ ES_HCAP (distance)
    uint32_t *hf;             // HEAP mode: heap entries ≥ ES_HCAP (slot | creation serial << 8)
    __device__ __forceinline__ explicit EsSlab(unsigned char *base)
    {
        plane = reinterpret_cast<double *>(base);
        base += static_cast<size_t>(ES_SLOTS) * 32;
        topo = reinterpret_cast<unsigned long long *>(base);
        base += static_cast<size_t>(ES_SLOTS) * 8;
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
#ifndef PK_EC_EXPERIMENT_NO_VAB // timing experiment only (results wrong): what do the pa / pb stores cost in L2 footprint
        double2 *v = reinterpret_cast<double2 *>(vab + 6 * i);
        v[0] = make_double2(s.pa.x, s.pa.y);
        v[1] = make_double2(s.pa.z, s.pb.x);
        v[2] = make_double2(s.pb.y, s.pb.z);
#endif
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
// per-pair input records (simplex, EpaInit) are read once: evict-first loads keep them from pushing the polytopes of
// the pairs in flight out of L2
template <class T> __device__ __forceinline__ T ec_ld_stream(const T *p)
{
#if defined(__CUDA_ARCH__) && defined(PK_EC_STREAM)
    return __ldcs(p);
#else
    return *p;
#endif
}
__device__ __forceinline__ void es_put_shape(EcSmem &sm, int which, const ShapeView &v, const BodyArrays &ba)
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
__device__ __forceinline__ ShapeView es_get_shape(const EcSmem &sm, int which, const BodyArrays &ba)
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

// HEAP mode: the first ES_HCAP heap entries live in the thread's shared-memory area (every sift starts there;
// polyhedron pairs rarely leave it), deeper ones in its slab.
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

// Everything a lane needs to start a pair except the simplex vertices, gathered by a kernel in which every lane
// works: the initial tetrahedron (build_initial_tetrahedron, collision.cpp:355-388: four face planes and the
// brute-force adjacency, ≈600 instructions) and the pair's key and contact slot, which the persistent kernel would
// otherwise reach through three dependent loads (simplex → pair → key / slot) every time a lane refills.
struct alignas(16) EpaInit
{
    double plane[4][4];         // normal xyz, distance of faces (0,1,2|3) (0,2,3|1) (0,3,1|2) (1,3,2|0)
    unsigned long long topo[4]; // vertices (flipped where the normal faced the opposite vertex), neighbours, serial
    uint64_t key;               // (body a << 32) | body b
    uint32_t out_slot;          // rank of the pair among the GJK hits = index of its contact record
    uint32_t flags;             // EPA_INIT_*
};
static_assert(sizeof(EpaInit) == 176, "EpaInit layout");
constexpr uint32_t EPA_INIT_BAD = 1u;    // a distance is NaN / inf: not for the recycling kernels
constexpr uint32_t EPA_INIT_PADDED = 2u; // simplex with fewer than four points: pad_simplex path, epa_kernel
constexpr uint32_t EPA_INIT_POLY = 4u;   // cost class 0 (no smooth shape): coplanar faces tie all the time → HEAP mode
constexpr uint32_t EC_META_HEAP = 1u << 31; // EcSmem::meta: the owner pops from the heap (S2 leaves the distance, not a key)

__global__ void __launch_bounds__(128)
epa_init_kernel(const SimplexRec *__restrict__ simplices, const unsigned long long *__restrict__ hit_count_ptr, uint64_t hit_capacity,
                const uint64_t *__restrict__ keys, const uint32_t *__restrict__ pair_a, const uint32_t *__restrict__ pair_b,
                const uint32_t *__restrict__ out_index, EpaInit *__restrict__ init)
{
    const uint64_t s = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    unsigned long long nhits = *hit_count_ptr;
    if (nhits > hit_capacity) nhits = hit_capacity;
    if (s >= nhits) return;
    const SimplexRec *r = simplices + s;
    EpaInit o;
    {
        const uint32_t pair = r->pair;
        uint32_t ia, ib;
        load_pair(keys, pair_a, pair_b, pair, ia, ib);
        o.key = (static_cast<uint64_t>(ia) << 32) | ib;
        o.out_slot = out_index[pair];
    }
    if ((r->n & 0xFFu) != 4u)
    {
        init[s].key = o.key;
        init[s].out_slot = o.out_slot;
        init[s].flags = EPA_INIT_PADDED;
        return;
    }
    d3 pv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        SupportPt sp;
        sp.pa = d3{r->v[i][0], r->v[i][1], r->v[i][2]};
        sp.pb = d3{r->v[i][3], r->v[i][4], r->v[i][5]};
        pv[i] = P(sp);
    }
    uint32_t tv[4]; // vertex triples, one byte each
    bool bad = false;
#pragma unroll
    for (int f = 0; f < 4; ++f)
    {
        d3 n;
        double dist;
        const bool flip = epa_face_plane(pv[fi4(f)], pv[fj4(f)], pv[fk4(f)], true, pv[fo4(f)], n, dist);
        tv[f] = static_cast<uint32_t>(fi4(f)) | (static_cast<uint32_t>(flip ? fk4(f) : fj4(f)) << 8) |
                (static_cast<uint32_t>(flip ? fj4(f) : fk4(f)) << 16);
        o.plane[f][0] = n.x;
        o.plane[f][1] = n.y;
        o.plane[f][2] = n.z;
        o.plane[f][3] = dist;
        if (!(fabs(dist) < 1e30)) bad = true; // NaN / inf: the key order would not be the heap's
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
    o.flags = (bad ? EPA_INIT_BAD : 0u) | (((r->n >> 8) & 0xFu) == 0u ? EPA_INIT_POLY : 0u);
    init[s] = o;
}

// 5 blocks of 64 threads per SM is what shared memory allows; the register file is split over the four schedulers of
// an SM (16 K registers each), so 10 warps need 3 warps per scheduler: at most 168 registers per thread
#ifndef PK_EC_MIN_BLOCKS
#define PK_EC_MIN_BLOCKS 5
#endif
#ifndef PK_EC_FETCH_MIN
#define PK_EC_FETCH_MIN 12
#endif
#ifdef PK_EC_STATS // host emulation only (tests/cpp): how often the rare paths run
extern unsigned long long g_ec_stats[16];
#define EC_STAT(i, v) __atomic_fetch_add(&g_ec_stats[i], static_cast<unsigned long long>(v), __ATOMIC_RELAXED)
#else
#define EC_STAT(i, v)
#endif

// One new face (collision.cpp:475-482 init_face + link_faces, 484-497 ring links) of the pair owned by lane
// `ot - wbase` of this warp, executed by whichever lane the edge was dealt to.
__device__ __forceinline__ void ec_make_face(EcSmem &shm, unsigned char *oslab, int ot, int e, int nh)
{
    const EsSlab osl(oslab);
    const uint32_t hc = shm.hz[e][ot];
    const int st = hz_start(hc), en = hz_end(hc), adj = hz_adj(hc), slot = hz_slot(hc);
    const uint32_t meta = shm.meta[ot];
    const d3 p{shm.pnew[0][ot], shm.pnew[1][ot], shm.pnew[2][ot]};
    const d3 cs = osl.vp(st), ce = osl.vp(en);
    const unsigned long long at = osl.topo[adj];
    d3 n;
    double dist;
    epa_face_plane(cs, ce, p, false, d3{0, 0, 0}, n, dist);
    osl.store_plane(slot, n, dist);
    bool bad = !(fabs(dist) < 1e30); // NaN / inf: the key order would not be the heap's
    if (meta & EC_META_HEAP)
    {
        if (e < ES_DNEW) shm.pop[ot].heap.dnew[e] = dist; // the owner pushes the faces in horizon order after S2
    }
    else if (slot < ES_KEYS)
        shm.pop[ot].key[slot] = __double2float_rd(dist);
    else
        osl.gkey[slot - ES_KEYS] = __double2float_rd(dist);
    // link_faces(f, adj_face, start, end): on the old face the shared edge starts at `end` (collision.cpp:305-313)
    const int e2 = (es_v(at, 0) == en) ? 0 : (es_v(at, 1) == en ? 1 : 2);
    if (es_v(at, e2) != en) bad = true; // unmatched link: slot recycling is no longer safe
    osl.set_adj(adj, e2, slot);
    // ring: the edge that starts at `end` gives the neighbour on edge 1, the edge that ends at `start` the one on
    // edge 2.  A proper horizon has exactly one of each for every edge, no self loop and no 2-cycle (which the
    // reference links one way only); anything else is handed back.
    int succ = 0xFF, pred = 0xFF, nsucc = 0, npred = 0;
    for (int j = 0; j < nh; ++j)
    {
        const uint32_t hj = shm.hz[j][ot];
        if (hz_start(hj) == en)
        {
            succ = hz_slot(hj);
            ++nsucc;
            if (hz_end(hj) == st) bad = true;
        }
        if (hz_end(hj) == st)
        {
            pred = hz_slot(hj);
            ++npred;
        }
    }
    if (nsucc != 1 || npred != 1 || st == en) bad = true;
    osl.topo[slot] = static_cast<unsigned long long>(st) | (static_cast<unsigned long long>(en) << 8) |
                     (static_cast<unsigned long long>(meta & 0xFFu) << 16) | (static_cast<unsigned long long>(adj) << 24) |
                     (static_cast<unsigned long long>(succ) << 32) | (static_cast<unsigned long long>(pred) << 40) |
                     (static_cast<unsigned long long>(((meta >> 8) & 0xFFFFu) + static_cast<uint32_t>(e)) << 48); // creation serial
    if (bad) shm.bad[ot] = 1u;
}

struct EcParams
{
    BodyArrays bodies;
    const SimplexRec *simplices;
    const unsigned long long *hit_count_ptr;
    uint64_t hit_capacity;
    const uint32_t *order;        // hit slots grouped by cost class: two smooth shapes, one, none
    ContactRec *contacts;
    uint8_t *valid;
    unsigned char *slabs;
    unsigned long long *cursor;   // into order[]
    unsigned long long *counters; // [0] valid contacts, [1] dropped: no room for the contact
    uint32_t *fallback;           // pairs handed to epa_kernel
    unsigned long long *fallback_count;
    unsigned long long *restart_count; // SCAN pairs started again in HEAP mode (statistics)
    const EpaInit *init;
    ContactRec *contacts_host;    // MIRROR: the caller's pinned buffer
};

// MIRROR: the pk_collide instance, which also delivers finished records to the caller's pinned buffer.
// HULLS: the instance for contexts that hold many-vertex hulls (support<true>, both support calls unrolled: the hull
// scans then keep their loads in flight; C4 31.5 against 33.8 ms) — sphere / box scenes run the other one, which does
// not carry that code (one copy of the support code in a two-trip loop: C3 12.6 against 13.5 ms).
template <bool MIRROR, bool HULLS>
__global__ void __launch_bounds__(ES_THREADS, PK_EC_MIN_BLOCKS) epa_coop_kernel(const __grid_constant__ EcParams P_)
{
    __shared__ EcSmem shm;
#ifdef PK_EC_TIMING
    unsigned long long ec_t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ec_t0));
#endif
    const BodyArrays &bodies = P_.bodies;
    const int t = threadIdx.x;
    const int lane = t & 31, wbase = t & ~31;
    const uint64_t tid = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    constexpr size_t slab_bytes = es_slab_bytes();
    unsigned char *const wslab = P_.slabs + (tid - static_cast<uint64_t>(lane)) * slab_bytes; // slab of lane 0 of this warp
    const EsSlab sl(wslab + static_cast<size_t>(lane) * slab_bytes);
    const EsHeapRef hp{&shm.pop[t], sl.hd, sl.hf};
    unsigned long long nhits = *P_.hit_count_ptr;
    if (nhits > P_.hit_capacity) nhits = P_.hit_capacity;
    const float INF = es_inf();

    bool active = false, done = false;
    bool heap = false;    // pop mode of the current pair
    bool swapped = false; // shape view 0 holds body b (see the fetch)
    bool restart = false; // the current pair met a tie only the heap can break: start it again in HEAP mode
    bool pending = false; // the pair has converged on face pend_face; its record is written at the next refill
    int pend_face = 0;
    int nverts = 0, iter = 0, hi = 0; // hi: slots [0, hi) have been used by the current polytope
    bool keys_dirty = true;           // the key area does not hold +inf beyond the current polytope
    int gdirty = ES_SLOTS;            // slab keys [ES_KEYS, gdirty) may hold something else than +inf
    unsigned long long fm0 = 0, fm1 = 0, fm2 = 0; // free slots, 64 per word
    uint32_t out_slot = 0, cur_sidx = 0;
    // tie breaking (see pop): lower bound of the distances of the lazily deleted heap entries the reference's
    // heap would still hold (faces killed by the flood fill; a pop removes every entry below the popped one)
    double stale_lb = 1e300;
    int batch_n = 0; // hz[0, batch_n) still describes the last batch of faces, in push order
    uint32_t n_valid = 0, n_dropped = 0;
    int heap_size = 0, nfaces = 0; // nfaces: faces created so far = serial of the next one
    constexpr unsigned FULL = 0xFFFFFFFFu;
    int fb = 0; // reason + 1 when the current pair has to be handed to epa_kernel
    bool flush_pending = false;
    uint32_t flush_slot = 0;

    auto is_free = [&](int f) -> bool
    {
        const unsigned long long w = (f < 64) ? fm0 : (f < 128 ? fm1 : fm2);
        return (w >> (f & 63)) & 1ull;
    };
    auto kill_slot = [&](int f)
    {
        if (!heap)
        {
            if (f < ES_KEYS)
                shm.pop[t].key[f] = INF;
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
    auto key_of = [&](int s) -> float { return (s < ES_KEYS) ? shm.pop[t].key[s] : sl.gkey[s - ES_KEYS]; };
    // slot of a new face.  SCAN: lowest free first (keeps the live polytope dense at the start of the slab and
    // inside the keyed slots).  HEAP: never-used slots first, so that as long as the polytope has created at most
    // ES_SLOTS faces a heap entry's slot equals its serial and "slot free" is the whole obsolete test; only bigger
    // polytopes recycle and have to compare serials.
    auto take_slot = [&]() -> int
    {
        int slot;
        if (heap && nfaces < ES_SLOTS)
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
            const unsigned long long i = atomicAdd(P_.fallback_count, 1ull);
            if (i < nhits) P_.fallback[i] = cur_sidx;
            fb = 0;
            active = false;
        }
        const unsigned m_active = __ballot_sync(FULL, active);
        const unsigned m_idle = __ballot_sync(FULL, !active && !done);
        if (m_active == 0 && m_idle == 0) break;
        const bool refill = !active && !done && (__popc(m_idle) >= PK_EC_FETCH_MIN || m_active == 0);
        if (refill && pending)
        {
            // collision.cpp:424-454 for the face the pair converged on (or the best guess after 64 iterations)
            es_write_result(sl, sl.load_plane(pend_face), sl.topo[pend_face], P_.contacts + out_slot, P_.init[cur_sidx].key,
                            MIRROR ? reinterpret_cast<ContactRec *>(&shm.pop[t]) : nullptr);
            P_.valid[out_slot] = 1;
            ++n_valid;
            pending = false;
            if constexpr (MIRROR)
            {
                flush_pending = true;
                flush_slot = out_slot;
                if (hi < 22) hi = 22; // the staged record covers the first 22 key slots: the next SCAN pair resets them
            }
        }
        if constexpr (MIRROR)
        {
            // pk_collide: finished records go to the caller's pinned buffer from here, so that no device→host copy
            // of the contacts has to wait for the kernel to end.  The record waits in the finishing lane's pop area;
            // eleven lanes store it with ONE 8-byte store each (88 contiguous bytes).  Letting the finishing lane
            // store to host memory itself — eleven dependent stores from one lane of a divergent warp — slowed the
            // kernel by more than the copy costs.
            unsigned m_flush = __ballot_sync(FULL, flush_pending);
            if (m_flush)
            {
                __syncwarp();
                while (m_flush)
                {
                    const int L = __ffs(static_cast<int>(m_flush)) - 1;
                    m_flush &= m_flush - 1u;
                    const uint32_t slot_l = __shfl_sync(FULL, flush_slot, L);
                    if (lane < 11)
                        reinterpret_cast<unsigned long long *>(P_.contacts_host + slot_l)[lane] =
                            reinterpret_cast<const unsigned long long *>(&shm.pop[wbase + L])[lane];
                }
                __syncwarp();
                flush_pending = false;
            }
        }
        if (refill || restart)
        {
            bool have = restart; // a restarted pair keeps its place, its slab and cur_sidx
            if (!have)
            {
                const unsigned long long slot = atomicAdd(P_.cursor, 1ull);
                if (slot >= nhits)
                    done = true;
                else
                {
                    cur_sidx = P_.order[slot];
                    have = true;
                }
            }
            if (have)
            {
                const EpaInit *in = P_.init + cur_sidx; // epa_init_kernel: tetrahedron, key, contact slot, cost class
                const uint4 meta4 = ec_ld_stream(reinterpret_cast<const uint4 *>(&in->key)); // key, contact slot, flags
                const uint32_t flags = meta4.w;
                const uint64_t key = (static_cast<uint64_t>(meta4.y) << 32) | meta4.x;
                out_slot = meta4.z;
                iter = 0;
                active = false;
                if (out_slot >= P_.hit_capacity)
                    ++n_dropped; // more GJK hits than contact records: the step reports PK_E_PAIR_OVERFLOW
                else if (flags & EPA_INIT_PADDED)
                    fb = 1; // pad_simplex path (collision.cpp:191-248): rare, left to epa_kernel
                else
                {
                    const bool was_heap = heap;
                    heap = restart || (flags & EPA_INIT_POLY);
                    const SimplexRec *r = P_.simplices + cur_sidx;
                    // The smoother shape goes into view 0 whichever body it is (sphere before many-vertex hull
                    // before the rest): a warp then runs one support path per view instead of two, with pairs listed
                    // by class.  `swapped` remembers which view is body a.
                    int rank_a = 0;
#pragma unroll 1
                    for (int w = 0; w < 2; ++w)
                    {
                        const ShapeView S = load_shape(bodies, w ? static_cast<uint32_t>(key & 0xFFFFFFFFu) : static_cast<uint32_t>(key >> 32));
                        const int rank = S.kind == KIND_SPHERE ? 0 : ((S.kind == KIND_HULL && S.nverts > HULL_PREFILTER_MIN) ? 1 : 2);
                        if (w == 0) rank_a = rank;
                        swapped = w == 1 && rank < rank_a;
                        es_put_shape(shm, w, S, bodies);
                    }
                    if (swapped)
                    {
                        const ShapeView S0 = es_get_shape(shm, 0, bodies), S1 = es_get_shape(shm, 1, bodies);
                        es_put_shape(shm, 0, S1, bodies);
                        es_put_shape(shm, 1, S0, bodies);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                    {
                        const double2 *rv = reinterpret_cast<const double2 *>(r->v[i]);
                        const double2 v0 = ec_ld_stream(rv), v1 = ec_ld_stream(rv + 1), v2 = ec_ld_stream(rv + 2);
                        SupportPt s;
                        s.pa = d3{v0.x, v0.y, v1.x};
                        s.pb = d3{v1.y, v2.x, v2.y};
                        sl.set_vert(i, s, P(s));
                    }
                    if (!heap)
                    {
                        // leftovers of the previous polytope; everything after a HEAP pair or a staged record, which
                        // occupied the same shared memory
                        const int dirty = (keys_dirty || was_heap) ? ES_KEYS : (hi < ES_KEYS ? hi : ES_KEYS);
                        for (int s = 4; s < dirty; ++s) shm.pop[t].key[s] = INF;
                        for (int s = ES_KEYS; s < gdirty; ++s) sl.gkey[s - ES_KEYS] = INF;
                        gdirty = ES_KEYS;
                        keys_dirty = false;
                    }
                    heap_size = 0;
#pragma unroll
                    for (int f = 0; f < 4; ++f)
                    {
                        const double2 *q = reinterpret_cast<const double2 *>(in->plane[f]);
                        const double2 n01 = ec_ld_stream(q), n2d = ec_ld_stream(q + 1);
                        sl.store_plane(f, d3{n01.x, n01.y, n2d.x}, n2d.y);
                        sl.topo[f] = ec_ld_stream(in->topo + f);
                        if (heap)
                        {
                            es_sift_up(hp, heap_size, n2d.y, static_cast<uint32_t>(f) | (static_cast<uint32_t>(f) << 8)); // push_face
                            ++heap_size;
                        }
                        else
                        {
                            shm.pop[t].key[f] = __double2float_rd(n2d.y);
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
                    if (flags & EPA_INIT_BAD) fb = 4;
                }
            }
            restart = false;
        }

        // ---------------- S1: the sequential part of one iteration, one lane per pair ----------------
        // Written as a sequence of stages, each under `go` and each closed before the next begins: a lane that leaves
        // the iteration (converged, handed back, restarted) clears `go`, and the lanes that go on meet again at the
        // next stage.  (With early exits out of nested blocks the compiler can only let them meet at the end of S1:
        // ncu showed the plane load after the pop at 14 of 32 lanes instead of 27.)
        int emit = 0; // horizon edges this lane hands to S2
        bool go = active && !fb;
        int min_face = -1;
        {
            {
                // ---- pop_face (collision.cpp:397-408) ----
                if (go && heap)
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
                else if (go)
                {
                    // minimal key: per 16-byte chunk the minimum of its four keys against the running minimum (no
                    // per-key index bookkeeping), then the winning chunk alone is looked at key by key.  Free slots
                    // hold +inf: they never win.  `more` counts the chunks that tie with the winner.
                    float m = INF;
                    int mc = -1, more = 0;
                    const int hs = hi < ES_KEYS ? hi : ES_KEYS;
                    const float4 *kq = reinterpret_cast<const float4 *>(shm.pop[t].key);
                    for (int c = 0; 4 * c < hs; ++c)
                    {
                        const float4 k4 = kq[c];
                        const float m4 = fminf(fminf(k4.x, k4.y), fminf(k4.z, k4.w));
                        const bool lt = m4 < m;
                        more = lt ? 0 : more + (m4 == m ? 1 : 0);
                        mc = lt ? c : mc;
                        m = lt ? m4 : m;
                    }
                    if (hi > ES_KEYS) // only polytopes past ≈45 iterations: the keys kept in the slab
                    {
                        const float4 *gq = reinterpret_cast<const float4 *>(sl.gkey);
                        for (int c = ES_KEYS / 4; 4 * c < hi; ++c)
                        {
                            const float4 k4 = gq[c - ES_KEYS / 4];
                            const float m4 = fminf(fminf(k4.x, k4.y), fminf(k4.z, k4.w));
                            const bool lt = m4 < m;
                            more = lt ? 0 : more + (m4 == m ? 1 : 0);
                            mc = lt ? c : mc;
                            m = lt ? m4 : m;
                        }
                    }
                    if (mc >= 0)
                    {
                        const float4 k4 = mc < ES_KEYS / 4 ? kq[mc] : reinterpret_cast<const float4 *>(sl.gkey)[mc - ES_KEYS / 4];
                        const int e0 = k4.x == m, e1 = k4.y == m, e2 = k4.z == m, e3 = k4.w == m;
                        min_face = 4 * mc + (e0 ? 0 : (e1 ? 1 : (e2 ? 2 : 3)));
                        more += e0 + e1 + e2 + e3 - 1;
                    }
                    EC_STAT(0, 1);
                    EC_STAT(1, hi);
                    if (min_face >= 0 && more > 0)
                    {
                        EC_STAT(2, 1);
                        // several live faces share the minimal float key (mirror-image faces of sphere–sphere
                        // polytopes differ in the last bits only): compare the exact distances
                        // (13 % of the pops of sphere–sphere pairs, 1.5 % for sphere–box: some lane of a warp is here
                        // in most iterations, so the candidates are collected first and their distances loaded together)
                        uint32_t cand = 0; // up to four more slots with the minimal key, one byte each, lowest first
                        int nc = 0;
                        for (int c = mc; 4 * c < hi; ++c)
                        {
                            const float4 k4 = c < ES_KEYS / 4 ? kq[c] : reinterpret_cast<const float4 *>(sl.gkey)[c - ES_KEYS / 4];
#ifdef PK_EC_CAND_SKIP // A/B: skipping chunks without a match up front measured no faster (12.86 vs 13.00 ms without)
                            if (!(k4.x == m || k4.y == m || k4.z == m || k4.w == m)) continue;
#endif
                            const float kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                            {
                                if (kk[u] == m && 4 * c + u != min_face)
                                {
                                    if (nc < 4) cand |= static_cast<uint32_t>(4 * c + u) << (8 * nc);
                                    ++nc;
                                }
                            }
                        }
                        double best = sl.plane[4 * min_face + 3];
                        bool tie = false;
                        if (nc <= 4)
                        {
                            double dd[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (i < nc) dd[i] = sl.plane[4 * static_cast<int>((cand >> (8 * i)) & 0xFFu) + 3];
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                            {
                                if (i >= nc) continue;
                                if (dd[i] < best)
                                {
                                    best = dd[i];
                                    min_face = static_cast<int>((cand >> (8 * i)) & 0xFFu);
                                    tie = false;
                                }
                                else if (dd[i] == best)
                                    tie = true;
                            }
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
                            EC_STAT(3, 1);
                            // Which of several equidistant faces std::pop_heap delivers depends on the heap's history.
                            // Two cases can be decided without it.  Let d be the tied minimum and suppose every
                            // lazily deleted entry still in the heap is farther than d (stale_lb).
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
                                restart = true; // same lane, same slab, from the simplex again, with the heap
                                active = false;
                                go = false;
                                atomicAdd(P_.restart_count, 1ull);
                            }
                        }
                    }
                }
                if (go && min_face < 0)
                {
                    P_.valid[out_slot] = 0; // heap exhausted → nullopt (collision.cpp:459,502)
                    active = false;
                    go = false;
                }
            }
            double4 mf = make_double4(0.0, 0.0, 0.0, 0.0);
            unsigned long long mt = 0;
            SupportPt sp{}; // minkowski_support (collision.h:41-49): pa = A.support(n), pb = B.support(−n)
            d3 p{0.0, 0.0, 0.0};
            if (go)
            {
                mf = sl.load_plane(min_face);
                mt = sl.topo[min_face];
                if (mf.w > stale_lb) stale_lb = mf.w; // entries closer than the popped face have left the heap
                const bool finished = iter >= 64;     // best guess after the loop (collision.cpp:500-503)
                if (!finished) ++iter;
#pragma unroll
                for (int k = 0; k < 3; ++k) // the flood fill starts with these: on their way during the support evaluation
                {
                    const int b = es_adj(mt, k);
                    if (b != 0xFF)
                    {
                        es_prefetch(sl.plane + 4 * b);
                        es_prefetch(sl.topo + b);
                    }
                }
                const d3 mn{mf.x, mf.y, mf.z};
                if (!finished)
                {
                    if constexpr (HULLS)
                    {
#pragma unroll
                        for (int w = 0; w < 2; ++w)
                        {
                            const bool is_b = (w == 1) != swapped; // view w holds body b
                            const ShapeView S = es_get_shape(shm, w, bodies);
                            const d3 q = support<true>(S, is_b ? -mn : mn);
                            if (is_b)
                                sp.pb = q;
                            else
                                sp.pa = q;
                        }
                    }
                    else
                    {
#pragma unroll 1
                        for (int w = 0; w < 2; ++w) // one copy of the support code (instruction cache)
                        {
                            const bool is_b = (w == 1) != swapped;
                            const ShapeView S = es_get_shape(shm, w, bodies);
#ifdef PK_EC_SUPPORT_BIG_ALWAYS
                            const d3 q = support<true>(S, is_b ? -mn : mn);
#else
                            const d3 q = support<false>(S, is_b ? -mn : mn);
#endif
                            if (is_b)
                                sp.pb = q;
                            else
                                sp.pa = q;
                        }
                    }
                }
                p = P(sp);
                if (finished || dot(mn, p) - mf.w < 1e-6) // converged (collision.cpp:465-466)
                {
                    pending = true;
                    pend_face = min_face;
                    active = false;
                    go = false;
                }
            }
            if (go)
            {
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
#ifdef PK_EC_FLOOD_PREFETCH
                        unsigned long long nt[3];
#endif
                        bool live[3];
#pragma unroll
                        for (int i = 0; i < 3; ++i)
                        {
                            const int a = es_adj(cur, i);
                            live[i] = a != 0xFF && !is_free(a);
                            if (live[i])
                            {
                                nf[i] = sl.load_plane(a);
#ifdef PK_EC_FLOOD_PREFETCH
                                nt[i] = sl.topo[a];
#endif
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
#ifdef PK_EC_FLOOD_PREFETCH // its neighbourhood will be needed when it is popped: measured slower in this kernel
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
#endif
                                }
                                else
                                    bad = true;
                            }
                            else
                            {
                                // horizon edge (cur.v[i], cur.v[i+1], a)
                                const int st = es_v(cur, i), en = es_v(cur, (i + 1) % 3);
                                if (nh < ES_HORIZON)
                                {
                                    shm.hz[nh][t] = static_cast<uint32_t>(st) | (static_cast<uint32_t>(en) << 7) | (static_cast<uint32_t>(a) << 14);
                                    ++nh;
                                }
                                else
                                    bad = true;
                            }
                        }
                        EC_STAT(4, 1);
                        if (depth == 0) break;
                        cur = sl.topo[static_cast<int>(stack & 0xFFull)];
                        stack >>= 8;
                        --depth;
                    }
                }
                batch_n = 0;
                const int nfree = __popcll(fm0) + __popcll(fm1) + __popcll(fm2);
                const bool full = nfree < nh || nverts >= ES_VERTS || (heap && heap_size + nh > ES_HEAP_MAX);
                if (nh == 0 && !bad)
                {
                    iter = 64; // empty horizon → best remaining face (collision.cpp:469,500-503)
                    go = false;
                }
                else if (bad || nh < 3 || full)
                {
                    fb = full ? 3 : 4;
                    go = false;
                }
                else
                    emit = nh;
            }
            if (go)
            {
                const int nh = emit;
                sl.set_vert(nverts, sp, p);
                shm.pnew[0][t] = p.x;
                shm.pnew[1][t] = p.y;
                shm.pnew[2][t] = p.z;
                shm.meta[t] = static_cast<uint32_t>(nverts) | (static_cast<uint32_t>(nfaces) << 8) | (heap ? EC_META_HEAP : 0u);
                shm.bad[t] = 0u;
                ++nverts;
                // slots of the new faces in horizon (= push) order.  (A shorter loop for SCAN lanes — clear the lowest set
                // bit of the first non-empty word — measured slower: EPA 12.9 → 14.2 ms.)
                for (int e = 0; e < nh; ++e)
                {
                    const int slot = take_slot();
                    if (slot >= hi) hi = slot + 1;
                    if (slot >= gdirty) gdirty = slot + 1;
                    ++nfaces;
                    shm.hz[e][t] |= static_cast<uint32_t>(slot) << 24;
                }
                EC_STAT(5, nh);
                EC_STAT(6, 1);
                EC_STAT(7 + (heap ? 1 : 0), 1);
            }
        }

        // ---------------- S2: the new faces of all pairs of the warp, dealt out evenly over its lanes ----------------
        int incl = emit;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const int v = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += v;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        if (total > 0)
        {
            __syncwarp(); // horizons, new vertices and slots are in shared memory / the slabs
            for (int base = 0; base < total; base += 32)
            {
                const int i = base + lane;
                int lo = 0; // number of lanes whose inclusive count is ≤ i = the owner of edge i
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1)
                {
                    const int v = __shfl_sync(FULL, incl, lo + step - 1);
                    if (v <= i) lo += step;
                }
                const int L = lo & 31;
                const int o_incl = __shfl_sync(FULL, incl, L);
                const int o_cnt = __shfl_sync(FULL, emit, L);
                if (i < total) ec_make_face(shm, wslab + static_cast<size_t>(L) * slab_bytes, wbase + L, i - (o_incl - o_cnt), o_cnt);
            }
            __syncwarp();
            if (emit)
            {
                if (shm.bad[t])
                    fb = 4;
                else
                {
                    if (heap)
                    {
                        const uint32_t serial0 = (shm.meta[t] >> 8) & 0xFFFFu;
                        for (int e = 0; e < emit; ++e) // push_face in horizon order (collision.cpp:479)
                        {
                            const int slot = hz_slot(shm.hz[e][t]);
                            const double dist = e < ES_DNEW ? shm.pop[t].heap.dnew[e] : sl.plane[4 * slot + 3];
                            es_sift_up(hp, heap_size, dist, static_cast<uint32_t>(slot) | ((serial0 + static_cast<uint32_t>(e)) << 8));
                            ++heap_size;
                        }
                    }
                    batch_n = emit;
                }
            }
        }
    }
    if (n_valid) atomicAdd(P_.counters + 0, static_cast<unsigned long long>(n_valid));
    if (n_dropped) atomicAdd(P_.counters + 1, static_cast<unsigned long long>(n_dropped));
#ifdef PK_EC_TIMING // debug build: when does each warp retire (how long is the tail of the persistent launch)
    if (lane == 0)
    {
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        g_ec_times[(tid / 32) & 8191][0] = ec_t0;
        g_ec_times[(tid / 32) & 8191][1] = now;
    }
#endif
}

} // namespace pk
