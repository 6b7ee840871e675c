// pk_narrowphase.cuh — batched GJK + EPA (reference src/collision.cpp:10-518), FP64, no FMA.
//
// K7a gjk_kernel   one thread per candidate pair: boolean GJK (collision.cpp:165-189).  Misses cost a
//                  couple of support evaluations and leave.  Hits append their terminating simplex to a
//                  dense hit list (warp-aggregated atomic) so that EPA sees only real work.
// K7b epa_kernel   persistent threads: every lane pulls the next hit from the list as soon as its
//                  previous pair converges (lanes never idle while work remains, which matters because
//                  EPA trip counts range from 1 to 64).  The polytope (faces / heap / vertices) lives
//                  in a per-thread slab of HBM laid out in 32-byte sectors: one face = one sector.
//
// Control flow and arithmetic follow the reference statement by statement so that results are
// bit-identical to the CPU oracle (oracle/pk_oracle.hpp); the face heap restates libstdc++'s
// __push_heap / __adjust_heap so that ties between equidistant faces resolve identically.
#pragma once

#include "pk_common.cuh"

namespace pk
{

constexpr int EPA_MAX_FACES = 768; // reference: unbounded (InlinedVector spills); overflow is flagged
constexpr int EPA_MAX_VERTS = 68;  // 4 + 64 iterations
#ifndef PK_EPA_MAX_HORIZON
#define PK_EPA_MAX_HORIZON 32
#endif
constexpr int EPA_MAX_HORIZON = PK_EPA_MAX_HORIZON; // reference InlinedVector<…,32> inline capacity; observed max 10
constexpr int EPA_MAX_STACK = 32;   // observed max 4
constexpr int EPA_THREADS = 64;

struct SupportPt
{
    d3 pa, pb; // p = pa - pb is recomputed where needed (bit-identical, saves a third of the state)
};
__device__ __forceinline__ d3 P(const SupportPt &s) { return s.pa - s.pb; }

// collision.h:41-49
template <bool BIG = true> __device__ __forceinline__ SupportPt minkowski_support(const ShapeView &a, const ShapeView &b, d3 d)
{
    SupportPt s;
    s.pa = support<BIG>(a, d);
    s.pb = support<BIG>(b, -d);
    return s;
}

// Terminating simplex handed from GJK to EPA: 208 bytes (16-byte aligned).
struct alignas(16) SimplexRec
{
    double v[4][6]; // pa xyz, pb xyz
    uint32_t n;
    uint32_t pair; // index into the pair list
};

// ---------------------------------------------------------------------------------------------
// GJK simplex handlers (collision.cpp:12-162).  Simplex order: newest point last.
// ---------------------------------------------------------------------------------------------
struct Simplex
{
    SupportPt pt[4];
    int n;
};

__device__ __forceinline__ void sx_erase(Simplex &s, int i)
{
    for (int k = i; k + 1 < s.n; ++k) s.pt[k] = s.pt[k + 1];
    --s.n;
}

// The handlers return the UN-normalised search direction; the caller normalises once.  Every
// branch of the reference ends in `direction = X.normalized()`, so this is the same arithmetic with
// one call site instead of eleven (smaller code: the r1 profile showed instruction-fetch stalls).

// collision.cpp:12-41
__device__ __forceinline__ d3 handle_line(Simplex &s)
{
    const d3 a = P(s.pt[1]);
    const d3 b = P(s.pt[0]);
    const d3 ab = b - a;
    const d3 ao = -a;
    if (dot(ab, ao) > 0.0)
    {
        d3 triple = cross(cross(ab, ao), ab);
        if (sqnorm(triple) < 1e-12)
        {
            d3 ab_hat = normalized(ab);
            d3 perp = cross(ab_hat, d3{0.0, 1.0, 0.0});
            if (sqnorm(perp) < 1e-12) perp = cross(ab_hat, d3{0.0, 0.0, 1.0});
            return perp;
        }
        return triple;
    }
    sx_erase(s, 0);
    return ao;
}

// collision.cpp:43-88
__device__ __forceinline__ d3 handle_triangle(Simplex &s)
{
    const d3 a = P(s.pt[2]);
    const d3 b = P(s.pt[1]);
    const d3 c = P(s.pt[0]);
    const d3 ab = b - a;
    const d3 ac = c - a;
    const d3 ao = -a;
    const d3 abc = cross(ab, ac);
    const d3 ab_perp = cross(ab, abc);
    if (dot(ab_perp, ao) > 0.0)
    {
        sx_erase(s, 0);
        d3 triple = cross(cross(ab, ao), ab);
        return (sqnorm(triple) < 1e-12) ? ao : triple;
    }
    const d3 ac_perp = cross(abc, ac);
    if (dot(ac_perp, ao) > 0.0)
    {
        sx_erase(s, 1);
        d3 triple = cross(cross(ac, ao), ac);
        return (sqnorm(triple) < 1e-12) ? ao : triple;
    }
    if (dot(abc, ao) <= 0.0)
    {
        SupportPt t = s.pt[0];
        s.pt[0] = s.pt[1];
        s.pt[1] = t;
        return -abc;
    }
    return abc;
}

// collision.cpp:90-147.  Returns true when the tetrahedron encloses the origin; otherwise the simplex
// has been rebuilt as the selected triangle ({c,b,a} / {d,c,a} / {b,d,a}) and the caller continues
// with handle_triangle, exactly as the reference's `return handle_triangle(simplex, direction)`.
// (The reference's `direction = abc.normalized()` before that call is dead: handle_triangle always
// overwrites the direction.)
__device__ __forceinline__ bool handle_tetrahedron(Simplex &s)
{
    const SupportPt sa = s.pt[3], sb = s.pt[2], sc = s.pt[1], sd = s.pt[0];
    const d3 a = P(sa), b = P(sb), c = P(sc), d = P(sd);
    const d3 ao = -a;
    d3 abc = cross(b - a, c - a);
    d3 acd = cross(c - a, d - a);
    d3 adb = cross(d - a, b - a);
    if (dot(abc, d - a) > 0.0) abc = -abc;
    if (dot(acd, b - a) > 0.0) acd = -acd;
    if (dot(adb, c - a) > 0.0) adb = -adb;
    int sel = -1;
    if (dot(abc, ao) > 0.0)
        sel = 0;
    else if (dot(acd, ao) > 0.0)
        sel = 1;
    else if (dot(adb, ao) > 0.0)
        sel = 2;
    if (sel < 0) return true;
    s.pt[0] = (sel == 0) ? sc : (sel == 1 ? sd : sb);
    s.pt[1] = (sel == 0) ? sb : (sel == 1 ? sc : sd);
    s.pt[2] = sa;
    s.n = 3;
    return false;
}

// One GJK iteration after the support point has been pushed (collision.cpp:149-162, 185).
// Returns true when the origin is enclosed.
__device__ __forceinline__ bool handle_simplex(Simplex &s, d3 &dir)
{
    if (s.n == 4 && handle_tetrahedron(s)) return true;
    d3 raw = (s.n == 3) ? handle_triangle(s) : handle_line(s);
    dir = normalized(raw);
    return false;
}

// collision.cpp:165-189.  Returns true with the terminating simplex when the shapes intersect.
// The first support (direction (1,0,0)) and the loop's supports share one call site.
__device__ __forceinline__ bool gjk_collision(const ShapeView &A, const ShapeView &B, Simplex &s)
{
    d3 dir{1.0, 0.0, 0.0};
    s.n = 0;
    for (int iter = -1; iter < 100; ++iter)
    {
        SupportPt np = minkowski_support(A, B, dir);
        d3 p = P(np);
        if (iter < 0)
        {
            s.pt[0] = np;
            s.n = 1;
            if (sqnorm(p) < 1e-12) return true;
            dir = -normalized(p);
            continue;
        }
        if (dot(p, dir) <= 0.0) return false;
        s.pt[s.n++] = np;
        if (handle_simplex(s, dir)) return true;
    }
    return false;
}

// What gjk_prefilter_kernel hands to gjk_kernel for a surviving pair: the first two support points, so that the
// survivor does not evaluate them a second time (half of the supports of a typical pair).  96 bytes per candidate
// pair, written for survivors only.
struct alignas(16) GjkCarry
{
    double v[2][6]; // pa xyz, pb xyz of simplex points 0 and 1; v[1][0] = NaN: the first point hit the origin, which
                    // gjk_collision answers with the one-point simplex (collision.cpp:174)
};
static_assert(sizeof(GjkCarry) == 96, "GjkCarry layout");

// gjk_collision (collision.cpp:165-189) from its third support on: the simplex holds the first two points, the
// second of which passed the separation test against direction −normalized(p0).  Written without early returns
// (one exit flag): lanes of a warp meet again at the head of every iteration.
template <bool BIG = true> __device__ __forceinline__ bool gjk_resume(const ShapeView &A, const ShapeView &B, Simplex &s)
{
    d3 dir{0.0, 0.0, 0.0};
    bool hit = handle_simplex(s, dir); // the line case of iteration 0
    bool alive = !hit;
    for (int iter = 1; alive && iter < 100; ++iter)
    {
        const SupportPt np = minkowski_support<BIG>(A, B, dir);
        if (dot(P(np), dir) <= 0.0)
            alive = false;
        else
        {
            s.pt[s.n++] = np;
            if (handle_simplex(s, dir))
            {
                hit = true;
                alive = false;
            }
        }
    }
    return hit;
}

// collision.cpp:191-248
__device__ __forceinline__ bool pad_simplex(const ShapeView &A, const ShapeView &B, Simplex &s)
{
    if (s.n == 1)
    {
        d3 dir{1.0, 0.0, 0.0};
        SupportPt p2 = minkowski_support(A, B, dir);
        if (sqnorm(P(p2) - P(s.pt[0])) < 1e-6) p2 = minkowski_support(A, B, -dir);
        s.pt[s.n++] = p2;
    }
    if (s.n == 2)
    {
        d3 line = P(s.pt[1]) - P(s.pt[0]);
        d3 dir = cross(normalized(line), d3{0.0, 1.0, 0.0});
        if (sqnorm(dir) < 1e-6) dir = cross(normalized(line), d3{0.0, 0.0, 1.0});
        dir = normalized(dir);
        SupportPt p3 = minkowski_support(A, B, dir);
        if (sqnorm(cross(line, P(p3) - P(s.pt[0]))) < 1e-6) p3 = minkowski_support(A, B, -dir);
        s.pt[s.n++] = p3;
    }
    if (s.n == 3)
    {
        d3 ab = P(s.pt[1]) - P(s.pt[0]);
        d3 ac = P(s.pt[2]) - P(s.pt[0]);
        d3 dir = normalized(cross(ab, ac));
        SupportPt p4 = minkowski_support(A, B, dir);
        if (fabs(dot(P(p4) - P(s.pt[0]), dir)) < 1e-6) p4 = minkowski_support(A, B, -dir);
        s.pt[s.n++] = p4;
    }
    d3 p3 = P(s.pt[3]);
    d3 ad = P(s.pt[0]) - p3;
    d3 bd = P(s.pt[1]) - p3;
    d3 cd = P(s.pt[2]) - p3;
    return fabs(dot(ad, cross(bd, cd))) > 1e-12;
}

// ---------------------------------------------------------------------------------------------
// K7a: GJK over the pair list.
//   pair source: either sorted u64 keys (a = key>>32, b = key&0xffffffff) or explicit a[]/b[] arrays.
//   Outputs: hit flag per pair, dense hit list with simplices.
// ---------------------------------------------------------------------------------------------

#ifndef PK_GJK_THREADS
#define PK_GJK_THREADS 128
#endif
#ifndef PK_GJK_MIN_BLOCKS
#define PK_GJK_MIN_BLOCKS 4
#endif

__device__ __forceinline__ void load_pair(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ pair_a,
                                          const uint32_t *__restrict__ pair_b, uint64_t k, uint32_t &ia, uint32_t &ib)
{
    if (keys)
    {
        uint64_t key = keys[k];
        ia = static_cast<uint32_t>(key >> 32);
        ib = static_cast<uint32_t>(key & 0xFFFFFFFFu);
    }
    else
    {
        ia = pair_a[k];
        ib = pair_b[k];
    }
}

constexpr uint32_t EPA_CLASSES = 9;
__device__ __forceinline__ uint32_t epa_cost_class(const ShapeView &A, const ShapeView &B)
{
    const bool big_a = A.kind == KIND_HULL && A.nverts > HULL_PREFILTER_MIN, big_b = B.kind == KIND_HULL && B.nverts > HULL_PREFILTER_MIN;
    const uint32_t smooth = (A.kind == KIND_SPHERE || big_a ? 1u : 0u) + (B.kind == KIND_SPHERE || big_b ? 1u : 0u);
    if (smooth == 0u) return 0u;
    const uint32_t nva = big_a ? A.nverts : 0u, nvb = big_b ? B.nverts : 0u;
    const uint32_t nv = nva > nvb ? nva : nvb;
    const uint32_t bucket = nv <= 48u ? 0u : (nv <= 96u ? 1u : (nv <= 192u ? 2u : 3u));
    return 1u + 4u * (smooth - 1u) + bucket;
}

// One thread per surviving pair.  (A persistent-lane variant with per-lane refill was measured in r1
// and lost: with a mean of 1.9 iterations per pair half the lanes refill every round and the set-up
// path — two gathered body loads — then sits on the critical path of every round.)
// CARRY: resume from the two support points gjk_prefilter_kernel<true> left; otherwise from the first support.
// BIG: the context holds hulls above HULL_PREFILTER_MIN vertices (support<BIG>).
template <bool CARRY, bool BIG = CARRY>
__global__ void __launch_bounds__(PK_GJK_THREADS, PK_GJK_MIN_BLOCKS)
gjk_kernel(BodyArrays bodies, const uint64_t *__restrict__ keys, const uint32_t *__restrict__ pair_a,
           const uint32_t *__restrict__ pair_b, const uint32_t *__restrict__ work, uint64_t work_stride,
           const unsigned long long *__restrict__ work_count /*[4]*/, uint8_t *__restrict__ hit,
           SimplexRec *__restrict__ simplices, unsigned long long *__restrict__ hit_count, uint64_t hit_capacity,
           unsigned long long *__restrict__ class_count /*[EPA_CLASSES]*/, const GjkCarry *__restrict__ carry /*[npairs]*/)
{
    uint64_t w = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    int c = 0; // class lists back to back: thread w works on entry w of their concatenation
    while (c < 4 && w >= work_count[c]) w -= work_count[c++];
    if (c == 4) return;
    const uint64_t k = work[c * work_stride + w];
    uint32_t ia, ib;
    load_pair(keys, pair_a, pair_b, k, ia, ib);
    ShapeView A = load_shape(bodies, ia);
    ShapeView B = load_shape(bodies, ib);
    Simplex s;
    bool h;
    {
        if constexpr (CARRY)
        {
            const double2 *cv = reinterpret_cast<const double2 *>(carry + k);
            const double2 c0 = cv[0], c1 = cv[1], c2 = cv[2], c3 = cv[3];
            s.pt[0].pa = d3{c0.x, c0.y, c1.x};
            s.pt[0].pb = d3{c1.y, c2.x, c2.y};
            s.n = 1;
            h = true;
            if (c3.x == c3.x)
            {
                const double2 c4 = cv[4], c5 = cv[5];
                s.pt[1].pa = d3{c3.x, c3.y, c4.x};
                s.pt[1].pb = d3{c4.y, c5.x, c5.y};
                s.n = 2;
                h = gjk_resume<true>(A, B, s);
            }
        }
        else
        {
            // gjk_collision from its first support (after gjk_filter_kernel; after gjk_prefilter_kernel<false>, where
            // evaluating two cheap supports again costs less than carrying 96 bytes per survivor through HBM)
            s.pt[0] = minkowski_support<BIG>(A, B, d3{1.0, 0.0, 0.0});
            s.n = 1;
            h = true;
            const d3 p0 = P(s.pt[0]);
            if (!(sqnorm(p0) < 1e-12))
            {
                const d3 dir = -normalized(p0);
                s.pt[1] = minkowski_support<BIG>(A, B, dir);
                s.n = 2;
                h = !(dot(P(s.pt[1]), dir) <= 0.0) && gjk_resume<BIG>(A, B, s); // collision.cpp:181-182
            }
        }
    }
    hit[k] = h ? 1 : 0;
    if (h)
    {
        unsigned long long slot = atomicAdd(hit_count, 1ull);
        if (slot < hit_capacity)
        {
            SimplexRec *r = simplices + slot;
            for (int i = 0; i < 4; ++i)
            {
                if (i < s.n)
                {
                    double2 *d = reinterpret_cast<double2 *>(&r->v[i][0]);
                    d[0] = make_double2(s.pt[i].pa.x, s.pt[i].pa.y);
                    d[1] = make_double2(s.pt[i].pa.z, s.pt[i].pb.x);
                    d[2] = make_double2(s.pt[i].pb.y, s.pt[i].pb.z);
                }
            }
            // EPA cost class: 0 = no "smooth" shape (spheres and many-vertex hulls need many more EPA iterations than
            // boxes, and their face distances practically never tie: epa_coop_kernel pops those heap-free, class 0
            // from the exact heap); 1-4 = one smooth shape, 5-8 = two, within each by the size of the largest hull
            // (spheres: none), so that the lanes of a warp scan hulls of similar length
            const uint32_t cls = epa_cost_class(A, B);
            {
                const unsigned peers = __match_any_sync(__activemask(), cls);
                if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(class_count + cls, static_cast<unsigned long long>(__popc(peers)));
            }
            r->n = static_cast<uint32_t>(s.n) | (cls << 8);
            r->pair = static_cast<uint32_t>(k);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// EPA polytope: per-thread slab in HBM + the hot top of the face heap in shared memory.
//
// The kernel is bound by the latency of dependent accesses to the polytope, not by bandwidth or FP64
// issue (ncu r1, 1 M bodies: 76 % of stall samples long_scoreboard, FP64 pipe 5 %, 41 GB of DRAM
// traffic for 0.4 GB of algorithmic bytes).  Layout and code are organised around that:
//   * the first EPA_HEAP_SMEM heap entries (levels 0-4 and part of 5: where every sift starts) live in
//     shared memory, transposed per thread; deeper entries spill to the slab
//   * heap entries carry a copy of the face distance, so sifting never touches the face array
//   * "obsolete" is a 768-bit per-thread bitset, so lazy heap deletion and the flood fill test it
//     without a memory round trip to the face
//   * face plane (normal, distance) is one 32-byte sector; topology (vertices, adjacency) 16 bytes
//   * the flood fill fetches the three neighbour planes of a face together
//   * new faces are linked in registers and stored once; ring links are found through a vertex→edge
//     table in O(h), with the reference's O(h²) loop kept for degenerate horizons
//   * vertex positions p = pa − pb are kept in their own 32-byte records (hot), pa / pb (only needed
//     for the final barycentric witness points) in cold ones
// ---------------------------------------------------------------------------------------------
struct alignas(16) FaceTopo
{
    uint16_t adj[3];
    uint8_t v[3];
    uint8_t _pad0;
    uint16_t _pad1[3];
};
static_assert(sizeof(FaceTopo) == 16, "FaceTopo is one 16-byte word");

struct alignas(16) HeapEnt
{
    double dist; // copy of face distance: distances never change after init_face
    uint32_t face;
    uint32_t _pad;
};

constexpr uint16_t EPA_NULL = 0xFFFFu;
#ifndef PK_EPA_HEAP_SMEM
#define PK_EPA_HEAP_SMEM 20
#endif
constexpr int EPA_HEAP_SMEM = PK_EPA_HEAP_SMEM;
constexpr int EPA_OBS_WORDS = EPA_MAX_FACES / 32;
constexpr size_t EPA_SLAB_BYTES = static_cast<size_t>(EPA_MAX_FACES) * (32 + sizeof(FaceTopo) + sizeof(HeapEnt)) +
                                  static_cast<size_t>(EPA_MAX_VERTS) * (32 + 48);

struct EpaSlab
{
    double4 *plane;   // normal xyz, distance (one sector per face)
    FaceTopo *topo;   // vertices, adjacency
    HeapEnt *heap;    // heap entries with index >= EPA_HEAP_SMEM
    double4 *vpos;    // p = pa - pb per polytope vertex (w unused)
    double *vab;      // pa xyz, pb xyz per polytope vertex
    __device__ __forceinline__ explicit EpaSlab(unsigned char *base)
    {
        plane = reinterpret_cast<double4 *>(base);
        base += static_cast<size_t>(EPA_MAX_FACES) * 32;
        topo = reinterpret_cast<FaceTopo *>(base);
        base += static_cast<size_t>(EPA_MAX_FACES) * sizeof(FaceTopo);
        heap = reinterpret_cast<HeapEnt *>(base);
        base += static_cast<size_t>(EPA_MAX_FACES) * sizeof(HeapEnt);
        vpos = reinterpret_cast<double4 *>(base);
        base += static_cast<size_t>(EPA_MAX_VERTS) * 32;
        vab = reinterpret_cast<double *>(base);
    }
    __device__ __forceinline__ d3 vp(int i) const
    {
        const double2 *v = reinterpret_cast<const double2 *>(vpos + i);
        double2 a = v[0], b = v[1];
        return {a.x, a.y, b.x};
    }
    __device__ __forceinline__ void set_vert(int i, const SupportPt &s, d3 p)
    {
        double2 *q = reinterpret_cast<double2 *>(vpos + i);
        q[0] = make_double2(p.x, p.y);
        q[1] = make_double2(p.z, 0.0);
        double2 *v = reinterpret_cast<double2 *>(vab + 6 * i);
        v[0] = make_double2(s.pa.x, s.pa.y);
        v[1] = make_double2(s.pa.z, s.pb.x);
        v[2] = make_double2(s.pb.y, s.pb.z);
    }
    __device__ __forceinline__ double4 load_plane(int f) const
    {
        const double2 *q = reinterpret_cast<const double2 *>(plane + f);
        double2 a = q[0], b = q[1];
        return make_double4(a.x, a.y, b.x, b.y);
    }
    __device__ __forceinline__ void store_plane(int f, d3 n, double dist)
    {
        double2 *q = reinterpret_cast<double2 *>(plane + f);
        q[0] = make_double2(n.x, n.y);
        q[1] = make_double2(n.z, dist);
    }
    __device__ __forceinline__ FaceTopo load_topo(int f) const
    {
        uint4 w = *reinterpret_cast<const uint4 *>(topo + f);
        FaceTopo t;
        t.adj[0] = static_cast<uint16_t>(w.x & 0xFFFFu);
        t.adj[1] = static_cast<uint16_t>(w.x >> 16);
        t.adj[2] = static_cast<uint16_t>(w.y & 0xFFFFu);
        t.v[0] = static_cast<uint8_t>((w.y >> 16) & 0xFFu);
        t.v[1] = static_cast<uint8_t>(w.y >> 24);
        t.v[2] = static_cast<uint8_t>(w.z & 0xFFu);
        return t;
    }
    __device__ __forceinline__ void store_topo(int f, const FaceTopo &t)
    {
        uint4 w;
        w.x = static_cast<uint32_t>(t.adj[0]) | (static_cast<uint32_t>(t.adj[1]) << 16);
        w.y = static_cast<uint32_t>(t.adj[2]) | (static_cast<uint32_t>(t.v[0]) << 16) | (static_cast<uint32_t>(t.v[1]) << 24);
        w.z = static_cast<uint32_t>(t.v[2]);
        w.w = 0;
        *reinterpret_cast<uint4 *>(topo + f) = w;
    }
    __device__ __forceinline__ void set_adj(int f, int e, uint16_t to) { topo[f].adj[e] = to; }
};

// Shared-memory part of the kernel state, transposed so that lane t owns column t.
struct EpaSmem
{
    double f[2][10][EPA_THREADS]; // shape views: p xyz, h xyz, q xyzw
    const double *verts[2][EPA_THREADS];
    const float4 *vf[2][EPA_THREADS];
    float hull_r[2][EPA_THREADS];
    int kind[2][EPA_THREADS];
    uint32_t nverts[2][EPA_THREADS];
    double hdist[EPA_HEAP_SMEM][EPA_THREADS];
    uint16_t hface[EPA_HEAP_SMEM][EPA_THREADS];
};

struct EpaHeap
{
    EpaSmem *sm;
    HeapEnt *g;
    __device__ __forceinline__ HeapEnt get(int k) const
    {
        HeapEnt e;
        if (k < EPA_HEAP_SMEM)
        {
            e.dist = sm->hdist[k][threadIdx.x];
            e.face = sm->hface[k][threadIdx.x];
            e._pad = 0;
        }
        else
            e = g[k];
        return e;
    }
    __device__ __forceinline__ double dist(int k) const { return (k < EPA_HEAP_SMEM) ? sm->hdist[k][threadIdx.x] : g[k].dist; }
    __device__ __forceinline__ void set(int k, const HeapEnt &e) const
    {
        if (k < EPA_HEAP_SMEM)
        {
            sm->hdist[k][threadIdx.x] = e.dist;
            sm->hface[k][threadIdx.x] = static_cast<uint16_t>(e.face);
        }
        else
            g[k] = e;
    }
};

// libstdc++ std::__push_heap with comp(a,b) = dist[a] > dist[b]  (collision.cpp:390-395)
__device__ __forceinline__ void heap_sift_up(const EpaHeap &h, int hole, const HeapEnt &value)
{
    int parent = (hole - 1) / 2;
    while (hole > 0)
    {
        HeapEnt pe = h.get(parent);
        if (!(pe.dist > value.dist)) break;
        h.set(hole, pe);
        hole = parent;
        parent = (hole - 1) / 2;
    }
    h.set(hole, value);
}
__device__ __forceinline__ void heap_push(const EpaHeap &h, int &size, uint32_t face, double dist)
{
    HeapEnt e;
    e.dist = dist;
    e.face = face;
    e._pad = 0;
    heap_sift_up(h, size, e);
    ++size;
}
// libstdc++ std::pop_heap (→ __pop_heap → __adjust_heap) followed by back()/pop_back()
__device__ __forceinline__ uint32_t heap_pop(const EpaHeap &h, int &size)
{
    if (size == 1)
    {
        size = 0;
        return h.get(0).face;
    }
    const int len = size - 1;
    const HeapEnt value = h.get(len);
    const uint32_t top_face = h.get(0).face;
    int hole = 0;
    int child = 0;
    while (child < (len - 1) / 2)
    {
        child = 2 * (child + 1);
        HeapEnt r = h.get(child), l = h.get(child - 1);
        if (r.dist > l.dist)
        {
            child--;
            h.set(hole, l);
        }
        else
            h.set(hole, r);
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2)
    {
        child = 2 * (child + 1);
        h.set(hole, h.get(child - 1));
        hole = child - 1;
    }
    heap_sift_up(h, hole, value); // __push_heap(first, hole, 0, value)
    size = len;
    return top_face;
}

// collision.cpp:273-297.  Computes normal / distance of face (i, j, k) from vertex positions already
// in registers; returns whether the orientation test flipped the face.
__device__ __forceinline__ bool epa_face_plane(d3 pi, d3 pj, d3 pk, bool has_opp, d3 popp, d3 &n, double &dist)
{
    d3 ab = pj - pi;
    d3 ac = pk - pi;
    n = cross(ab, ac);
    if (sqnorm(n) < 1e-12)
        n = d3{0.0, 0.0, 0.0};
    else
        n = normalized(n);
    bool flip = false;
    if (has_opp && dot(n, popp - pi) > 0.0)
    {
        flip = true;
        n = -n;
    }
    dist = dot(n, pi);
    return flip;
}

// collision.cpp:424-454
__device__ __forceinline__ void epa_write_result(const EpaSlab &sl, double4 nd, const FaceTopo &t, ContactRec *out, uint64_t key, ContactRec *mirror = nullptr)
{
    d3 n{nd.x, nd.y, nd.z};
    const double2 *q0 = reinterpret_cast<const double2 *>(sl.vab + 6 * t.v[0]);
    const double2 *q1 = reinterpret_cast<const double2 *>(sl.vab + 6 * t.v[1]);
    const double2 *q2 = reinterpret_cast<const double2 *>(sl.vab + 6 * t.v[2]);
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
    out->key = key; // 88-byte records are only 8-byte aligned: scalar stores
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
    if (mirror) *mirror = *out; // pk_collide: the caller's pinned copy (this kernel only sees the rare hand-backs)
}

__device__ __forceinline__ void smem_put_shape(EpaSmem &sm, int which, const ShapeView &v)
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
__device__ __forceinline__ ShapeView smem_get_shape(const EpaSmem &sm, int which)
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
__device__ __forceinline__ SupportPt minkowski_support_smem(const EpaSmem &sm, d3 d)
{
    SupportPt s;
    {
        ShapeView A = smem_get_shape(sm, 0);
        s.pa = support(A, d);
    }
    {
        ShapeView B = smem_get_shape(sm, 1);
        s.pb = support(B, -d);
    }
    return s;
}

// Hit slots grouped by EPA cost class (heaviest first) so that the lanes of a warp work on pairs of
// similar shape: similar trip counts, same support code path.  order[rank] = slot.
__global__ void __launch_bounds__(256)
epa_order_kernel(const SimplexRec *__restrict__ simplices, const unsigned long long *__restrict__ hit_count_ptr,
                 uint64_t hit_capacity, const unsigned long long *__restrict__ class_count,
                 unsigned long long *__restrict__ class_fill, uint32_t *__restrict__ order)
{
    unsigned long long nhits = *hit_count_ptr;
    if (nhits > hit_capacity) nhits = hit_capacity;
    __shared__ unsigned long long base[EPA_CLASSES]; // heaviest class first
    if (threadIdx.x == 0)
    {
        unsigned long long run = 0;
        for (int c = static_cast<int>(EPA_CLASSES) - 1; c >= 0; --c)
        {
            base[c] = run;
            run += class_count[c];
        }
    }
    __syncthreads();
    for (unsigned long long s = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; s < nhits;
         s += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
    {
        const uint32_t cls = (simplices[s].n >> 8) & 0xFu;
        // one atomic per (warp, class) instead of one per hit: a few hot addresses would serialise
        const unsigned peers = __match_any_sync(__activemask(), cls);
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(peers) - 1;
        unsigned long long first = 0;
        if (lane == leader) first = atomicAdd(class_fill + cls, static_cast<unsigned long long>(__popc(peers)));
        first = __shfl_sync(peers, first, leader);
        const unsigned long long pos = base[cls] + first + __popc(peers & ((1u << lane) - 1u));
        if (pos < hit_capacity) order[pos] = static_cast<uint32_t>(s);
    }
}

#ifndef PK_EPA_MIN_BLOCKS
#define PK_EPA_MIN_BLOCKS 6
#endif
#ifndef PK_EPA_FETCH_MIN
#define PK_EPA_FETCH_MIN 6
#endif

// ---------------------------------------------------------------------------------------------
// K7b: persistent EPA (collision.cpp:411-504).
//   out_index[pair] gives the slot of the pair's contact in the key-sorted contact array
//   (exclusive scan of the GJK hit flags); valid[slot] = 1 when EPA produced a value.
//   Lanes whose pair has finished wait until PK_EPA_FETCH_MIN lanes of the warp are idle (or nobody
//   is left running) before the set-up path runs, so that path is never executed for one lane.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EPA_THREADS, PK_EPA_MIN_BLOCKS)
epa_kernel(BodyArrays bodies, const uint64_t *__restrict__ keys, const uint32_t *__restrict__ pair_a,
           const uint32_t *__restrict__ pair_b, const SimplexRec *__restrict__ simplices,
           const unsigned long long *__restrict__ hit_count_ptr, uint64_t hit_capacity,
           const uint32_t *__restrict__ out_index, const uint32_t *__restrict__ order, ContactRec *__restrict__ contacts,
           uint8_t *__restrict__ valid, unsigned char *__restrict__ slabs, unsigned long long *__restrict__ cursor,
           unsigned long long *__restrict__ counters /* [0]=valid contacts, [1]=overflow */, ContactRec *contacts_host = nullptr)
{
    __shared__ EpaSmem shm;
    const uint64_t tid = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    EpaSlab sl(slabs + tid * EPA_SLAB_BYTES);
    EpaHeap heap{&shm, sl.heap};
    unsigned long long nhits = *hit_count_ptr;
    if (nhits > hit_capacity) nhits = hit_capacity;

    bool active = false, done = false;
    int nfaces = 0, nverts = 0, heap_size = 0, iter = 0;
    uint32_t out_slot = 0;
    uint64_t key = 0;
    unsigned long long n_valid = 0, n_over = 0;
    uint32_t obs[EPA_OBS_WORDS];           // obsolete flags of the current polytope
    uint8_t edge_of_start[EPA_MAX_VERTS];  // vertex → horizon edge starting there (0xFF = none)
    uint8_t edge_of_end[EPA_MAX_VERTS];
    for (int i = 0; i < EPA_MAX_VERTS; ++i) edge_of_start[i] = edge_of_end[i] = 0xFF;
    constexpr unsigned FULL = 0xFFFFFFFFu;

    for (;;)
    {
        const unsigned m_active = __ballot_sync(FULL, active);
        const unsigned m_idle = __ballot_sync(FULL, !active && !done);
        if (m_active == 0 && m_idle == 0) break;
        if (!active && !done && (__popc(m_idle) >= PK_EPA_FETCH_MIN || m_active == 0))
        {
            unsigned long long slot = atomicAdd(cursor, 1ull);
            if (slot >= nhits)
                done = true;
            else
            {
                const SimplexRec *r = simplices + order[slot];
                uint32_t pair = r->pair;
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
                if (out_slot >= hit_capacity)
                {
                    ++n_over; // more GJK hits than contact records (the step reports PK_E_PAIR_OVERFLOW): nothing to write to
                    continue;
                }
                Simplex s;
                s.n = static_cast<int>(r->n & 0xFFu);
                for (int i = 0; i < 4; ++i)
                {
                    if (i < s.n)
                    {
                        s.pt[i].pa = d3{r->v[i][0], r->v[i][1], r->v[i][2]};
                        s.pt[i].pb = d3{r->v[i][3], r->v[i][4], r->v[i][5]};
                    }
                }
                bool ok = true;
                {
                    ShapeView A = load_shape(bodies, ia);
                    ShapeView B = load_shape(bodies, ib);
                    smem_put_shape(shm, 0, A);
                    smem_put_shape(shm, 1, B);
                    if (s.n < 4) ok = pad_simplex(A, B, s);
                }
                if (!ok)
                    valid[out_slot] = 0; // degenerate: treated as no collision (collision.cpp:414-415)
                else
                {
                    d3 pv[4];
                    for (int i = 0; i < 4; ++i)
                    {
                        pv[i] = P(s.pt[i]);
                        sl.set_vert(i, s.pt[i], pv[i]);
                    }
                    nverts = 4;
                    nfaces = 4;
                    heap_size = 0;
                    for (int w = 0; w < EPA_OBS_WORDS; ++w) obs[w] = 0u;
                    // build_initial_tetrahedron (collision.cpp:355-388): faces (0,1,2|3) (0,2,3|1) (0,3,1|2) (1,3,2|0)
                    const int fi[4] = {0, 0, 0, 1}, fj[4] = {1, 2, 3, 3}, fk[4] = {2, 3, 1, 2}, fo[4] = {3, 1, 2, 0};
                    FaceTopo tl[4];
                    for (int f = 0; f < 4; ++f)
                    {
                        d3 n;
                        double dist;
                        bool flip = epa_face_plane(pv[fi[f]], pv[fj[f]], pv[fk[f]], true, pv[fo[f]], n, dist);
                        tl[f].v[0] = static_cast<uint8_t>(fi[f]);
                        tl[f].v[1] = static_cast<uint8_t>(flip ? fk[f] : fj[f]);
                        tl[f].v[2] = static_cast<uint8_t>(flip ? fj[f] : fk[f]);
                        tl[f].adj[0] = tl[f].adj[1] = tl[f].adj[2] = EPA_NULL;
                        sl.store_plane(f, n, dist);
                        heap_push(heap, heap_size, static_cast<uint32_t>(f), dist);
                    }
                    for (int i = 0; i < 4; ++i)
                        for (int j = i + 1; j < 4; ++j)
                            for (int e1 = 0; e1 < 3; ++e1)
                            {
                                uint8_t u1 = tl[i].v[e1], v1 = tl[i].v[(e1 + 1) % 3];
                                for (int e2 = 0; e2 < 3; ++e2)
                                {
                                    uint8_t u2 = tl[j].v[e2], v2 = tl[j].v[(e2 + 1) % 3];
                                    if (u1 == v2 && v1 == u2)
                                    {
                                        tl[i].adj[e1] = static_cast<uint16_t>(j);
                                        tl[j].adj[e2] = static_cast<uint16_t>(i);
                                    }
                                }
                            }
                    for (int f = 0; f < 4; ++f) sl.store_topo(f, tl[f]);
                    iter = 0;
                    active = true;
                }
            }
        }
        if (!active) continue;

        // ---- one EPA iteration, or the post-loop "best guess" when iter == 64 -----------------
        // pop_face(): skip obsolete entries (collision.cpp:397-408)
        int min_face = -1;
        while (heap_size > 0)
        {
            uint32_t f = heap_pop(heap, heap_size);
            if (!((obs[f >> 5] >> (f & 31u)) & 1u))
            {
                min_face = static_cast<int>(f);
                break;
            }
        }
        if (min_face < 0)
        {
            valid[out_slot] = 0; // heap exhausted → nullopt (collision.cpp:459,502)
            active = false;
            continue;
        }
        const double4 mf = sl.load_plane(min_face);
        const FaceTopo mt = sl.load_topo(min_face);
        if (iter >= 64)
        {
            epa_write_result(sl, mf, mt, contacts + out_slot, key, contacts_host ? contacts_host + out_slot : nullptr); // best guess (collision.cpp:500-503)
            valid[out_slot] = 1;
            ++n_valid;
            active = false;
            continue;
        }
        ++iter;
        d3 mn{mf.x, mf.y, mf.z};
        SupportPt sp = minkowski_support_smem(shm, mn);
        d3 p = P(sp);
        if (dot(mn, p) - mf.w < 1e-6)
        {
            epa_write_result(sl, mf, mt, contacts + out_slot, key, contacts_host ? contacts_host + out_slot : nullptr); // converged (collision.cpp:465-466)
            valid[out_slot] = 1;
            ++n_valid;
            active = false;
            continue;
        }

        // find_silhouette (collision.cpp:315-353): DFS, LIFO stack, edge order preserved.
        uint16_t stack[EPA_MAX_STACK];
        uint8_t hz_start[EPA_MAX_HORIZON], hz_end[EPA_MAX_HORIZON];
        uint16_t hz_adj[EPA_MAX_HORIZON];
        int sp_top = 0, nh = 0;
        bool overflow = false;
        obs[min_face >> 5] |= 1u << (min_face & 31);
        {
            FaceTopo cf = mt;
            for (;;)
            {
                double4 nf[3];
                bool live[3];
#pragma unroll
                for (int i = 0; i < 3; ++i)
                {
                    uint16_t nidx = cf.adj[i];
                    live[i] = nidx != EPA_NULL && !((obs[nidx >> 5] >> (nidx & 31u)) & 1u);
                    if (live[i]) nf[i] = sl.load_plane(nidx);
                }
#pragma unroll
                for (int i = 0; i < 3; ++i)
                {
                    if (!live[i]) continue;
                    uint16_t nidx = cf.adj[i];
                    // a neighbour reached through two edges of this face: the second visit must see the
                    // flag set by the first one
                    if ((obs[nidx >> 5] >> (nidx & 31u)) & 1u) continue;
                    if (dot(d3{nf[i].x, nf[i].y, nf[i].z}, p) > nf[i].w + 1e-6)
                    {
                        obs[nidx >> 5] |= 1u << (nidx & 31);
                        if (sp_top < EPA_MAX_STACK)
                            stack[sp_top++] = nidx;
                        else
                            overflow = true;
                    }
                    else
                    {
                        if (nh < EPA_MAX_HORIZON)
                        {
                            hz_start[nh] = cf.v[i];
                            hz_end[nh] = cf.v[(i + 1) % 3];
                            hz_adj[nh] = nidx;
                            ++nh;
                        }
                        else
                            overflow = true;
                    }
                }
                if (sp_top == 0) break;
                cf = sl.load_topo(stack[--sp_top]);
            }
        }
        if (nh == 0 && !overflow)
        {
            // empty horizon → leave the loop and return the best remaining face (collision.cpp:469,500-503)
            iter = 64;
            continue;
        }
        if (overflow || nfaces + nh > EPA_MAX_FACES || nverts >= EPA_MAX_VERTS)
        {
            valid[out_slot] = 0;
            ++n_over;
            active = false;
            continue;
        }
        sl.set_vert(nverts, sp, p);
        const int p_idx = nverts++;
        const int first_new = nfaces;
        nfaces += nh;
        // new faces (start, end, p_idx), no orientation flip (collision.cpp:475-482)
        double new_dist[EPA_MAX_HORIZON];
        for (int e = 0; e < nh; ++e)
        {
            d3 n;
            double dist;
            epa_face_plane(sl.vp(hz_start[e]), sl.vp(hz_end[e]), p, false, d3{0, 0, 0}, n, dist);
            new_dist[e] = dist;
            sl.store_plane(first_new + e, n, dist);
            // link_faces(f, adj_face, start, end): on the old face the shared edge starts at `end`
            FaceTopo b = sl.load_topo(hz_adj[e]);
            int e2 = (b.v[0] == hz_end[e]) ? 0 : (b.v[1] == hz_end[e] ? 1 : 2);
            sl.set_adj(hz_adj[e], e2, static_cast<uint16_t>(first_new + e));
        }
        // ring links among the new faces (collision.cpp:484-497).  Face e = (start_e, end_e, p_idx);
        // link_faces(i, j, end_i, p_idx) sets adj[edge of end_i in face i] = j and adj[edge of p_idx in
        // face j] = i.  For a proper horizon (every vertex starts exactly one edge and ends exactly one,
        // no self loop, no 2-cycle) every successor link is made exactly once whatever the visiting
        // order, so a vertex→edge table gives the same result in O(h).  Anything else takes the
        // reference's i<j double loop with its overwrite order.
        {
            uint16_t nadj[EPA_MAX_HORIZON][3];
            bool proper = nh >= 3;
            for (int e = 0; e < nh; ++e)
            {
                nadj[e][0] = hz_adj[e]; // link_faces(f, adj, start, end) on the new face: edge of `start` (= 0)
                nadj[e][1] = EPA_NULL;
                nadj[e][2] = EPA_NULL;
                if (edge_of_start[hz_start[e]] != 0xFF || edge_of_end[hz_end[e]] != 0xFF || hz_start[e] == hz_end[e]) proper = false;
                edge_of_start[hz_start[e]] = static_cast<uint8_t>(e);
                edge_of_end[hz_end[e]] = static_cast<uint8_t>(e);
            }
            if (proper)
            {
                for (int e = 0; e < nh; ++e)
                {
                    const uint8_t j = edge_of_start[hz_end[e]];
                    if (j == 0xFF) continue;
                    if (edge_of_start[hz_end[j]] == e) proper = false; // 2-cycle: reference links only one way
                }
            }
            if (proper)
            {
                for (int e = 0; e < nh; ++e)
                {
                    const uint8_t j = edge_of_start[hz_end[e]];
                    if (j == 0xFF) continue;
                    nadj[e][1] = static_cast<uint16_t>(first_new + j);
                    nadj[j][2] = static_cast<uint16_t>(first_new + e);
                }
            }
            else
            {
                for (int i = 0; i < nh; ++i)
                    for (int j = i + 1; j < nh; ++j)
                    {
                        int a, b, va;
                        if (hz_end[i] == hz_start[j])
                        {
                            a = i;
                            b = j;
                            va = hz_end[i];
                        }
                        else if (hz_start[i] == hz_end[j])
                        {
                            a = j;
                            b = i;
                            va = hz_end[j];
                        }
                        else
                            continue;
                        int e1 = (hz_start[a] == va) ? 0 : (hz_end[a] == va ? 1 : 2);
                        int e2 = (hz_start[b] == p_idx) ? 0 : (hz_end[b] == p_idx ? 1 : 2);
                        nadj[a][e1] = static_cast<uint16_t>(first_new + b);
                        nadj[b][e2] = static_cast<uint16_t>(first_new + a);
                    }
            }
            for (int e = 0; e < nh; ++e)
            {
                edge_of_start[hz_start[e]] = 0xFF;
                edge_of_end[hz_end[e]] = 0xFF;
                FaceTopo t;
                t.adj[0] = nadj[e][0];
                t.adj[1] = nadj[e][1];
                t.adj[2] = nadj[e][2];
                t.v[0] = hz_start[e];
                t.v[1] = hz_end[e];
                t.v[2] = static_cast<uint8_t>(p_idx);
                sl.store_topo(first_new + e, t);
            }
        }
        for (int e = 0; e < nh; ++e) heap_push(heap, heap_size, static_cast<uint32_t>(first_new + e), new_dist[e]);
    }
    if (n_valid) atomicAdd(counters + 0, n_valid);
    if (n_over) atomicAdd(counters + 1, n_over);
}

// contact_point(info, a, b) (collision_phases.h:78-82): the witness points of every contact in the bodies' own
// frames, particle::project_to_local = orientation.conjugate() * (world_point - pos) (core/particle.h:107-108).
// This is what narrow_phase::calculate stores in its manifolds; computed on request, one thread per contact.
struct ContactPointRec
{
    double local_a[3];
    double local_b[3];
};
__global__ void __launch_bounds__(256)
contact_points_kernel(const ContactRec *__restrict__ contacts, uint64_t n, const double *__restrict__ pos,
                      const double *__restrict__ quat, ContactPointRec *__restrict__ out)
{
    const uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (k >= n) return;
    const ContactRec c = contacts[k];
    const uint32_t ia = static_cast<uint32_t>(c.key >> 32), ib = static_cast<uint32_t>(c.key & 0xFFFFFFFFu);
    const dq qa{quat[4ull * ia], quat[4ull * ia + 1], quat[4ull * ia + 2], quat[4ull * ia + 3]};
    const dq qb{quat[4ull * ib], quat[4ull * ib + 1], quat[4ull * ib + 2], quat[4ull * ib + 3]};
    const d3 la = rotate(conjugate(qa), d3{c.world_a[0], c.world_a[1], c.world_a[2]} - d3{pos[3ull * ia], pos[3ull * ia + 1], pos[3ull * ia + 2]});
    const d3 lb = rotate(conjugate(qb), d3{c.world_b[0], c.world_b[1], c.world_b[2]} - d3{pos[3ull * ib], pos[3ull * ib + 1], pos[3ull * ib + 2]});
    ContactPointRec r;
    r.local_a[0] = la.x; r.local_a[1] = la.y; r.local_a[2] = la.z;
    r.local_b[0] = lb.x; r.local_b[1] = lb.y; r.local_b[2] = lb.z;
    out[k] = r;
}

// pk_gjk_epa_batch: one record per requested pair (zeros + hit = 0 for misses).
__global__ void expand_contacts_kernel(const uint8_t *__restrict__ hit, const uint32_t *__restrict__ index,
                                       const uint8_t *__restrict__ valid, const ContactRec *__restrict__ compact,
                                       const uint32_t *__restrict__ pa, const uint32_t *__restrict__ pb, uint64_t n,
                                       uint64_t capacity, ContactRec *__restrict__ out, uint8_t *__restrict__ out_hit)
{
    uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (k >= n) return;
    bool h = hit[k] != 0;
    uint32_t slot = index[k];
    if (h) h = slot < capacity && valid[slot] != 0; // hits beyond the contact capacity have no record

    ContactRec r;
    if (h)
        r = compact[slot];
    else
    {
        r.key = (static_cast<uint64_t>(pa[k]) << 32) | pb[k];
        for (int j = 0; j < 3; ++j) r.normal[j] = r.world_a[j] = r.world_b[j] = 0.0;
        r.depth = 0.0;
    }
    out[k] = r;
    out_hit[k] = h ? 1 : 0;
}

} // namespace pk
