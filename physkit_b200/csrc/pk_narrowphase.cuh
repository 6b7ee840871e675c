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
constexpr int EPA_MAX_HORIZON = 64;
constexpr int EPA_MAX_STACK = 64;
constexpr int EPA_THREADS = 64;

struct SupportPt
{
    d3 pa, pb; // p = pa - pb is recomputed where needed (bit-identical, saves a third of the state)
};
__device__ __forceinline__ d3 P(const SupportPt &s) { return s.pa - s.pb; }

// collision.h:41-49
__device__ __forceinline__ SupportPt minkowski_support(const ShapeView &a, const ShapeView &b, d3 d)
{
    SupportPt s;
    s.pa = support(a, d);
    s.pb = support(b, -d);
    return s;
}

// Terminating simplex handed from GJK to EPA: 200 bytes.
struct SimplexRec
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

// collision.cpp:12-41
__device__ __forceinline__ void handle_line(Simplex &s, d3 &dir)
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
            dir = normalized(perp);
        }
        else
            dir = normalized(triple);
    }
    else
    {
        sx_erase(s, 0);
        dir = normalized(ao);
    }
}

// collision.cpp:43-88
__device__ __forceinline__ void handle_triangle(Simplex &s, d3 &dir)
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
        dir = (sqnorm(triple) < 1e-12) ? normalized(ao) : normalized(triple);
        return;
    }
    const d3 ac_perp = cross(abc, ac);
    if (dot(ac_perp, ao) > 0.0)
    {
        sx_erase(s, 1);
        d3 triple = cross(cross(ac, ao), ac);
        dir = (sqnorm(triple) < 1e-12) ? normalized(ao) : normalized(triple);
        return;
    }
    if (dot(abc, ao) <= 0.0)
    {
        SupportPt t = s.pt[0];
        s.pt[0] = s.pt[1];
        s.pt[1] = t;
        dir = normalized(-abc);
    }
    else
        dir = normalized(abc);
}

// collision.cpp:90-147; returns true when the tetrahedron encloses the origin
__device__ __forceinline__ bool handle_tetrahedron(Simplex &s, d3 &dir)
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
    if (dot(abc, ao) > 0.0)
    {
        s.pt[0] = sc;
        s.pt[1] = sb;
        s.pt[2] = sa;
        s.n = 3;
        handle_triangle(s, dir);
        return false;
    }
    if (dot(acd, ao) > 0.0)
    {
        s.pt[0] = sd;
        s.pt[1] = sc;
        s.pt[2] = sa;
        s.n = 3;
        handle_triangle(s, dir);
        return false;
    }
    if (dot(adb, ao) > 0.0)
    {
        s.pt[0] = sb;
        s.pt[1] = sd;
        s.pt[2] = sa;
        s.n = 3;
        handle_triangle(s, dir);
        return false;
    }
    return true;
}

// collision.cpp:165-189.  Returns true with the terminating simplex when the shapes intersect.
__device__ __forceinline__ bool gjk_collision(const ShapeView &A, const ShapeView &B, Simplex &s)
{
    d3 dir{1.0, 0.0, 0.0};
    s.pt[0] = minkowski_support(A, B, dir);
    s.n = 1;
    d3 p0 = P(s.pt[0]);
    if (sqnorm(p0) < 1e-12) return true;
    dir = -normalized(p0);
    for (int iter = 0; iter < 100; ++iter)
    {
        SupportPt np = minkowski_support(A, B, dir);
        if (dot(P(np), dir) <= 0.0) return false;
        s.pt[s.n++] = np;
        if (s.n == 2)
            handle_line(s, dir);
        else if (s.n == 3)
            handle_triangle(s, dir);
        else if (handle_tetrahedron(s, dir))
            return true;
    }
    return false;
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
struct BodyArrays
{
    const ShapeRec *shapes;
    const double *verts;
    const double *pos;
    const double *quat;
    const uint32_t *shape_id;
};

__global__ void __launch_bounds__(128)
gjk_kernel(BodyArrays bodies, const uint64_t *__restrict__ keys, const uint32_t *__restrict__ pair_a,
           const uint32_t *__restrict__ pair_b, uint64_t npairs, uint8_t *__restrict__ hit,
           SimplexRec *__restrict__ simplices, unsigned long long *__restrict__ hit_count, uint64_t hit_capacity)
{
    uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (k >= npairs) return;
    uint32_t ia, ib;
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
    ShapeView A = load_shape(bodies.shapes, bodies.verts, bodies.pos, bodies.quat, bodies.shape_id, ia);
    ShapeView B = load_shape(bodies.shapes, bodies.verts, bodies.pos, bodies.quat, bodies.shape_id, ib);
    Simplex s;
    bool h = gjk_collision(A, B, s);
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
                    r->v[i][0] = s.pt[i].pa.x;
                    r->v[i][1] = s.pt[i].pa.y;
                    r->v[i][2] = s.pt[i].pa.z;
                    r->v[i][3] = s.pt[i].pb.x;
                    r->v[i][4] = s.pt[i].pb.y;
                    r->v[i][5] = s.pt[i].pb.z;
                }
            }
            r->n = static_cast<uint32_t>(s.n);
            r->pair = static_cast<uint32_t>(k);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// EPA polytope in a per-thread HBM slab.
// ---------------------------------------------------------------------------------------------
struct alignas(16) FaceTopo
{
    uint16_t adj[3];
    uint16_t _pad0;
    uint8_t v[3];
    uint8_t obsolete;
    uint32_t _pad1;
};
static_assert(sizeof(FaceTopo) == 16, "FaceTopo is half a sector");

struct alignas(16) HeapEnt
{
    double dist; // copy of face distance: distances never change after init_face
    uint32_t face;
    uint32_t _pad;
};

constexpr uint16_t EPA_NULL = 0xFFFFu;
constexpr size_t EPA_SLAB_BYTES = static_cast<size_t>(EPA_MAX_FACES) * (32 + sizeof(FaceTopo) + sizeof(HeapEnt)) +
                                  static_cast<size_t>(EPA_MAX_VERTS) * 48;

struct EpaSlab
{
    double4 *fnd;    // normal xyz, distance
    FaceTopo *topo;  // vertices, adjacency, obsolete
    HeapEnt *heap;   // binary heap, min distance at the root
    double *verts;   // pa xyz, pb xyz per polytope vertex
    __device__ __forceinline__ explicit EpaSlab(unsigned char *base)
    {
        fnd = reinterpret_cast<double4 *>(base);
        topo = reinterpret_cast<FaceTopo *>(base + static_cast<size_t>(EPA_MAX_FACES) * 32);
        heap = reinterpret_cast<HeapEnt *>(base + static_cast<size_t>(EPA_MAX_FACES) * (32 + sizeof(FaceTopo)));
        verts = reinterpret_cast<double *>(base + static_cast<size_t>(EPA_MAX_FACES) * (32 + sizeof(FaceTopo) + sizeof(HeapEnt)));
    }
    __device__ __forceinline__ d3 vp(int i) const
    {
        const double2 *v = reinterpret_cast<const double2 *>(verts + 6 * i);
        double2 a = v[0], b = v[1], c = v[2];
        return {a.x - b.y, a.y - c.x, b.x - c.y}; // pa - pb
    }
    __device__ __forceinline__ void set_vert(int i, const SupportPt &s)
    {
        double2 *v = reinterpret_cast<double2 *>(verts + 6 * i);
        v[0] = make_double2(s.pa.x, s.pa.y);
        v[1] = make_double2(s.pa.z, s.pb.x);
        v[2] = make_double2(s.pb.y, s.pb.z);
    }
};

// libstdc++ std::__push_heap with comp(a,b) = dist[a] > dist[b]  (collision.cpp:390-395)
__device__ __forceinline__ void heap_sift_up(HeapEnt *h, int hole, int top, HeapEnt value)
{
    int parent = (hole - 1) / 2;
    while (hole > top && h[parent].dist > value.dist)
    {
        h[hole] = h[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    h[hole] = value;
}
__device__ __forceinline__ void heap_push(HeapEnt *h, int &size, uint32_t face, double dist)
{
    HeapEnt e;
    e.dist = dist;
    e.face = face;
    e._pad = 0;
    heap_sift_up(h, size, 0, e);
    ++size;
}
// libstdc++ std::pop_heap (→ __pop_heap → __adjust_heap) followed by back()/pop_back()
__device__ __forceinline__ uint32_t heap_pop(HeapEnt *h, int &size)
{
    if (size == 1)
    {
        size = 0;
        return h[0].face;
    }
    const int len = size - 1;
    HeapEnt value = h[len];
    uint32_t top_face = h[0].face;
    int hole = 0;
    int child = 0;
    while (child < (len - 1) / 2)
    {
        child = 2 * (child + 1);
        if (h[child].dist > h[child - 1].dist) child--;
        h[hole] = h[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2)
    {
        child = 2 * (child + 1);
        h[hole] = h[child - 1];
        hole = child - 1;
    }
    heap_sift_up(h, hole, 0, value);
    size = len;
    return top_face;
}

// collision.cpp:273-297
__device__ __forceinline__ double epa_init_face(EpaSlab &sl, int f, int i, int j, int k, int opposite)
{
    d3 pi = sl.vp(i);
    d3 ab = sl.vp(j) - pi;
    d3 ac = sl.vp(k) - pi;
    d3 n = cross(ab, ac);
    if (sqnorm(n) < 1e-12)
        n = d3{0.0, 0.0, 0.0};
    else
        n = normalized(n);
    uint8_t v1 = static_cast<uint8_t>(j), v2 = static_cast<uint8_t>(k);
    if (opposite >= 0 && dot(n, sl.vp(opposite) - pi) > 0.0)
    {
        uint8_t t = v1;
        v1 = v2;
        v2 = t;
        n = -n;
    }
    double dist = dot(n, pi);
    sl.fnd[f] = make_double4(n.x, n.y, n.z, dist);
    FaceTopo t;
    t.adj[0] = t.adj[1] = t.adj[2] = EPA_NULL;
    t._pad0 = 0;
    t.v[0] = static_cast<uint8_t>(i);
    t.v[1] = v1;
    t.v[2] = v2;
    t.obsolete = 0;
    t._pad1 = 0;
    sl.topo[f] = t;
    return dist;
}

// collision.cpp:305-313
__device__ __forceinline__ void epa_link(EpaSlab &sl, int f1, int f2, int va, int vb)
{
    FaceTopo a = sl.topo[f1];
    int e1 = (a.v[0] == va) ? 0 : (a.v[1] == va ? 1 : 2);
    sl.topo[f1].adj[e1] = static_cast<uint16_t>(f2);
    FaceTopo b = sl.topo[f2];
    int e2 = (b.v[0] == vb) ? 0 : (b.v[1] == vb ? 1 : 2);
    sl.topo[f2].adj[e2] = static_cast<uint16_t>(f1);
}

// collision.cpp:424-454
__device__ __forceinline__ void epa_write_result(const EpaSlab &sl, int f, ContactRec *out, uint64_t key)
{
    double4 nd = sl.fnd[f];
    FaceTopo t = sl.topo[f];
    d3 n{nd.x, nd.y, nd.z};
    const double *q0 = sl.verts + 6 * t.v[0], *q1 = sl.verts + 6 * t.v[1], *q2 = sl.verts + 6 * t.v[2];
    d3 a0{q0[0], q0[1], q0[2]}, b0{q0[3], q0[4], q0[5]};
    d3 a1{q1[0], q1[1], q1[2]}, b1{q1[3], q1[4], q1[5]};
    d3 a2{q2[0], q2[1], q2[2]}, b2{q2[3], q2[4], q2[5]};
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

// ---------------------------------------------------------------------------------------------
// K7b: persistent EPA (collision.cpp:411-504).
//   out_index[pair] gives the slot of the pair's contact in the key-sorted contact array
//   (exclusive scan of the GJK hit flags); valid[slot] = 1 when EPA produced a value.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EPA_THREADS)
epa_kernel(BodyArrays bodies, const uint64_t *__restrict__ keys, const uint32_t *__restrict__ pair_a,
           const uint32_t *__restrict__ pair_b, const SimplexRec *__restrict__ simplices,
           const unsigned long long *__restrict__ hit_count_ptr, uint64_t hit_capacity,
           const uint32_t *__restrict__ out_index, ContactRec *__restrict__ contacts, uint8_t *__restrict__ valid,
           unsigned char *__restrict__ slabs, unsigned long long *__restrict__ cursor,
           unsigned long long *__restrict__ counters /* [0]=valid contacts, [1]=overflow */)
{
    const uint64_t tid = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    EpaSlab sl(slabs + tid * EPA_SLAB_BYTES);
    unsigned long long nhits = *hit_count_ptr;
    if (nhits > hit_capacity) nhits = hit_capacity;

    bool active = false;
    ShapeView A, B;
    int nfaces = 0, nverts = 0, heap_size = 0, iter = 0;
    uint32_t out_slot = 0;
    uint64_t key = 0;
    unsigned long long n_valid = 0, n_over = 0;

    for (;;)
    {
        if (!active)
        {
            unsigned long long slot = atomicAdd(cursor, 1ull);
            if (slot >= nhits) break;
            const SimplexRec *r = simplices + slot;
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
            A = load_shape(bodies.shapes, bodies.verts, bodies.pos, bodies.quat, bodies.shape_id, ia);
            B = load_shape(bodies.shapes, bodies.verts, bodies.pos, bodies.quat, bodies.shape_id, ib);
            Simplex s;
            s.n = static_cast<int>(r->n);
            for (int i = 0; i < 4; ++i)
            {
                if (i < s.n)
                {
                    s.pt[i].pa = d3{r->v[i][0], r->v[i][1], r->v[i][2]};
                    s.pt[i].pb = d3{r->v[i][3], r->v[i][4], r->v[i][5]};
                }
            }
            if (s.n < 4 && !pad_simplex(A, B, s))
            {
                valid[out_slot] = 0; // degenerate: treated as no collision (collision.cpp:414-415)
                continue;
            }
            for (int i = 0; i < 4; ++i) sl.set_vert(i, s.pt[i]);
            nverts = 4;
            nfaces = 4;
            heap_size = 0;
            // build_initial_tetrahedron (collision.cpp:355-388)
            heap_push(sl.heap, heap_size, 0, epa_init_face(sl, 0, 0, 1, 2, 3));
            heap_push(sl.heap, heap_size, 1, epa_init_face(sl, 1, 0, 2, 3, 1));
            heap_push(sl.heap, heap_size, 2, epa_init_face(sl, 2, 0, 3, 1, 2));
            heap_push(sl.heap, heap_size, 3, epa_init_face(sl, 3, 1, 3, 2, 0));
            for (int i = 0; i < 4; ++i)
                for (int j = i + 1; j < 4; ++j)
                {
                    FaceTopo fi = sl.topo[i];
                    FaceTopo fj = sl.topo[j];
                    for (int e1 = 0; e1 < 3; ++e1)
                    {
                        uint8_t u1 = fi.v[e1], v1 = fi.v[(e1 + 1) % 3];
                        for (int e2 = 0; e2 < 3; ++e2)
                        {
                            uint8_t u2 = fj.v[e2], v2 = fj.v[(e2 + 1) % 3];
                            if (u1 == v2 && v1 == u2)
                            {
                                fi.adj[e1] = static_cast<uint16_t>(j);
                                fj.adj[e2] = static_cast<uint16_t>(i);
                            }
                        }
                    }
                    sl.topo[i] = fi;
                    sl.topo[j] = fj;
                }
            iter = 0;
            active = true;
        }

        // ---- one EPA iteration, or the post-loop "best guess" when iter == 64 -----------------
        // pop_face(): skip obsolete entries (collision.cpp:397-408)
        int min_face = -1;
        while (heap_size > 0)
        {
            uint32_t f = heap_pop(sl.heap, heap_size);
            if (!sl.topo[f].obsolete)
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
        if (iter >= 64)
        {
            epa_write_result(sl, min_face, contacts + out_slot, key); // best guess (collision.cpp:500-503)
            valid[out_slot] = 1;
            ++n_valid;
            active = false;
            continue;
        }
        ++iter;
        double4 mf = sl.fnd[min_face];
        d3 mn{mf.x, mf.y, mf.z};
        SupportPt sp = minkowski_support(A, B, mn);
        d3 p = P(sp);
        if (dot(mn, p) - mf.w < 1e-6)
        {
            epa_write_result(sl, min_face, contacts + out_slot, key); // converged (collision.cpp:465-466)
            valid[out_slot] = 1;
            ++n_valid;
            active = false;
            continue;
        }

        // find_silhouette (collision.cpp:315-353): DFS, LIFO stack, edge order preserved
        uint16_t stack[EPA_MAX_STACK];
        uint8_t hz_start[EPA_MAX_HORIZON], hz_end[EPA_MAX_HORIZON];
        uint16_t hz_adj[EPA_MAX_HORIZON];
        int sp_top = 0, nh = 0;
        bool overflow = false;
        stack[sp_top++] = static_cast<uint16_t>(min_face);
        sl.topo[min_face].obsolete = 1;
        while (sp_top > 0)
        {
            int cur = stack[--sp_top];
            FaceTopo cf = sl.topo[cur];
            for (int i = 0; i < 3; ++i)
            {
                uint16_t nidx = cf.adj[i];
                if (nidx == EPA_NULL) continue;
                if (sl.topo[nidx].obsolete) continue;
                double4 nf = sl.fnd[nidx];
                if (dot(d3{nf.x, nf.y, nf.z}, p) > nf.w + 1e-6)
                {
                    sl.topo[nidx].obsolete = 1;
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
        sl.set_vert(nverts, sp);
        const int p_idx = nverts++;
        const int first_new = nfaces;
        for (int e = 0; e < nh; ++e)
        {
            int f = nfaces++;
            double dist = epa_init_face(sl, f, hz_start[e], hz_end[e], p_idx, -1);
            epa_link(sl, f, hz_adj[e], hz_start[e], hz_end[e]);
            heap_push(sl.heap, heap_size, static_cast<uint32_t>(f), dist);
        }
        for (int i = 0; i < nh; ++i)
            for (int j = i + 1; j < nh; ++j)
            {
                if (hz_end[i] == hz_start[j])
                    epa_link(sl, first_new + i, first_new + j, hz_end[i], p_idx);
                else if (hz_start[i] == hz_end[j])
                    epa_link(sl, first_new + j, first_new + i, hz_end[j], p_idx);
            }
    }
    if (n_valid) atomicAdd(counters + 0, n_valid);
    if (n_over) atomicAdd(counters + 1, n_over);
}

// pk_gjk_epa_batch: one record per requested pair (zeros + hit = 0 for misses).
__global__ void expand_contacts_kernel(const uint8_t *__restrict__ hit, const uint32_t *__restrict__ index,
                                       const uint8_t *__restrict__ valid, const ContactRec *__restrict__ compact,
                                       const uint32_t *__restrict__ pa, const uint32_t *__restrict__ pb, uint64_t n,
                                       ContactRec *__restrict__ out, uint8_t *__restrict__ out_hit)
{
    uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (k >= n) return;
    bool h = hit[k] != 0;
    uint32_t slot = index[k];
    if (h) h = valid[slot] != 0;
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
