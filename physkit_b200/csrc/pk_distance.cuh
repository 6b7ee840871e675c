// pk_distance.cuh — batched GJK closest-distance query (pk_gjk_distance_batch).
//
// BASELINE.json's north_star names "batched GJK intersection/distance".  The reference has the intersection half only
// (gjk_collision answers yes / no, src/collision.cpp:165-189, and a separated pair is std::nullopt to gjk_epa,
// :511-518): there is no reference arithmetic to reproduce here, so this is the textbook distance form of GJK over
// the SAME support mappings the boolean query uses (support<BIG>, pk_common.cuh = bounds.h:164-174, :539-548,
// :328-329, src/mesh.cpp:341-358), checked against a brute-force distance over all feature pairs on the oracle side
// and against separation certificates (tests/test_gjk_distance_host.py, tests/test_gpu_distance.py).
//
// One thread per pair.  v = the point of the current simplex of A ⊖ B closest to the origin, w = the support point of
// A ⊖ B along −v; v·w / |v| is a lower bound of the distance, |v| an upper bound; the iteration stops when they meet
// to 1e-12 of |v|, when the support point adds nothing new (polytopes: the exact optimum is reached after finitely
// many steps) or when rounding stops the descent.  The closest point of a simplex is found by Voronoi regions
// (vertex, edge, face), which also yields the barycentric weights of the witness points a = Σ λ·pa, b = Σ λ·pb.
// A sphere enters as its centre with a margin of r: a ball's support points never repeat, so the plain iteration
// would only converge in the limit, while centre distance − r is exact.
//
// FP64, no FMA contraction (the library is built with -fmad=false): the host build of this file in the test tree
// (g++ -ffp-contract=off) produces the same bits, which is how the kernel's logic is tested without a GPU.
#pragma once

#include "pk_narrowphase.cuh"

namespace pk
{

// 64-byte result record, mirrors pk_distance of include/pk_collide.h
struct DistanceRec
{
    uint64_t key;      // (a << 32) | b as given
    double distance;   // > 0 for separated shapes, 0 when they touch or overlap
    double point_a[3]; // closest point of a (world frame); zeros when not separated
    double point_b[3]; // closest point of b
};
static_assert(sizeof(DistanceRec) == 64, "DistanceRec must match pk_distance");

constexpr int DIST_MAX_ITERS = 64;
constexpr double DIST_REL_GAP = 1e-12; // stop when |v|² − v·w ≤ this · |v|²

struct DistSimplex
{
    SupportPt pt[4];
    double lam[4];
    int n;
};

// closest point of segment ab to the origin: weights of a and b (zero = the vertex leaves the simplex)
__device__ __forceinline__ void dist_segment(d3 a, d3 b, double &la, double &lb)
{
    const d3 ab = b - a;
    const double t = -dot(a, ab);
    if (!(t > 0.0))
    {
        la = 1.0;
        lb = 0.0;
        return;
    }
    const double den = sqnorm(ab);
    if (!(t < den))
    {
        la = 0.0;
        lb = 1.0;
        return;
    }
    lb = t / den;
    la = 1.0 - lb;
}

// closest point of triangle abc to the origin by Voronoi regions (vertex regions first, then edges, then the face)
__device__ __forceinline__ void dist_triangle(d3 a, d3 b, d3 c, double &la, double &lb, double &lc)
{
    const d3 ab = b - a, ac = c - a;
    const double d1 = -dot(ab, a), d2 = -dot(ac, a);
    la = lb = lc = 0.0;
    if (d1 <= 0.0 && d2 <= 0.0)
    {
        la = 1.0;
        return;
    }
    const double d3_ = -dot(ab, b), d4 = -dot(ac, b);
    if (d3_ >= 0.0 && d4 <= d3_)
    {
        lb = 1.0;
        return;
    }
    const double vc = d1 * d4 - d3_ * d2;
    if (vc <= 0.0 && d1 >= 0.0 && d3_ <= 0.0)
    {
        const double v = d1 / (d1 - d3_);
        la = 1.0 - v;
        lb = v;
        return;
    }
    const double d5 = -dot(ab, c), d6 = -dot(ac, c);
    if (d6 >= 0.0 && d5 <= d6)
    {
        lc = 1.0;
        return;
    }
    const double vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0)
    {
        const double w = d2 / (d2 - d6);
        la = 1.0 - w;
        lc = w;
        return;
    }
    const double va = d3_ * d6 - d5 * d4;
    if (va <= 0.0 && (d4 - d3_) >= 0.0 && (d5 - d6) >= 0.0)
    {
        const double w = (d4 - d3_) / ((d4 - d3_) + (d5 - d6));
        lb = 1.0 - w;
        lc = w;
        return;
    }
    const double sum = (va + vb) + vc;
    lb = vb / sum;
    lc = vc / sum;
    la = (1.0 - lb) - lc;
}

// Is the origin on the other side of plane abc than d?  A flat tetrahedron (d in the plane to within 1e-10 of its own
// size: the sign of the volume is rounding noise) counts as "outside" for every face: its hull is the union of its
// faces, and the closest of them is taken.
__device__ __forceinline__ bool dist_outside(d3 a, d3 b, d3 c, d3 d)
{
    const d3 n = cross(b - a, c - a);
    const d3 ad = d - a;
    const double so = -dot(a, n), sd = dot(ad, n);
    const bool flat = !(sd * sd > 1e-20 * (sqnorm(n) * sqnorm(ad)));
    return flat || so * sd < 0.0;
}

__device__ __forceinline__ d3 dist_point(const DistSimplex &s)
{
    d3 v{0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (i < s.n) v = v + s.lam[i] * P(s.pt[i]);
    return v;
}

// Replace the simplex by the smallest sub-simplex that holds its point closest to the origin, with the weights of
// that point.  → DIST_INSIDE when the origin lies inside a tetrahedron (the shapes intersect), DIST_STALLED when
// rounding left no vertex with a positive weight (the caller keeps its previous point).
constexpr int DIST_OK = 0, DIST_INSIDE = 1, DIST_STALLED = 2;
__device__ __forceinline__ int dist_reduce(DistSimplex &s)
{
    double l[4] = {0.0, 0.0, 0.0, 0.0};
    if (s.n == 1)
        l[0] = 1.0;
    else if (s.n == 2)
        dist_segment(P(s.pt[0]), P(s.pt[1]), l[0], l[1]);
    else if (s.n == 3)
        dist_triangle(P(s.pt[0]), P(s.pt[1]), P(s.pt[2]), l[0], l[1], l[2]);
    else
    {
        const d3 w0 = P(s.pt[0]), w1 = P(s.pt[1]), w2 = P(s.pt[2]), w3 = P(s.pt[3]);
        double best = 1e308;
        bool any = false;
        // face (i,j,k) with the fourth vertex o on its far side
        auto face = [&](d3 a, d3 b, d3 c, d3 o, int i, int j, int k)
        {
            if (!dist_outside(a, b, c, o)) return;
            double la, lb, lc;
            dist_triangle(a, b, c, la, lb, lc);
            const d3 q = (la * a + lb * b) + lc * c;
            const double sq = sqnorm(q);
            if (!any || sq < best)
            {
                any = true;
                best = sq;
#pragma unroll
                for (int m = 0; m < 4; ++m) l[m] = m == i ? la : (m == j ? lb : (m == k ? lc : 0.0));
            }
        };
        face(w0, w1, w2, w3, 0, 1, 2);
        face(w0, w2, w3, w1, 0, 2, 3);
        face(w0, w3, w1, w2, 0, 3, 1);
        face(w1, w3, w2, w0, 1, 3, 2);
        if (!any) return DIST_INSIDE;
    }
    // keep the vertices with a positive weight, order preserved
    int n = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        if (i < s.n && l[i] > 0.0)
        {
#pragma unroll
            for (int m = 0; m < 4; ++m)
                if (m == n && m <= i)
                {
                    s.pt[m] = s.pt[i];
                    s.lam[m] = l[i];
                }
            ++n;
        }
    }
    s.n = n;
    return n > 0 ? DIST_OK : DIST_STALLED;
}

// support mapping of the shape's core: a sphere is its centre (+ margin r), everything else itself
template <bool BIG> __device__ __forceinline__ d3 dist_core_support(const ShapeView &s, d3 d)
{
    if (s.kind == KIND_SPHERE) return s.p;
    return support<BIG>(s, d);
}
__device__ __forceinline__ double dist_margin(const ShapeView &s) { return s.kind == KIND_SPHERE ? s.h.x : 0.0; }

// → true when the shapes are separated; then distance > 0 and pa / pb are the closest points
template <bool BIG> __device__ __forceinline__ bool gjk_distance_pair(const ShapeView &A_, const ShapeView &B_, double &distance, d3 &out_a, d3 &out_b)
{
    // The iteration runs in a frame whose origin is A's reference point: support points are centre + offset, and with
    // centres of 1e4 next to offsets of 1e-3 the offsets would lose the digits the distance is made of.  The
    // difference of the two centres is computed once (exact when they are within a factor of two of each other).
    ShapeView A = A_, B = B_;
    const d3 O = A_.p;
    A.p = d3{0.0, 0.0, 0.0};
    B.p = B_.p - O;
    if (A.kind == KIND_AABB) A.h = A_.h - O; // world box: p = min, h = max
    if (B.kind == KIND_AABB) B.h = B_.h - O;
    DistSimplex s;
    {
        const d3 d0{1.0, 0.0, 0.0};
        s.pt[0].pa = dist_core_support<BIG>(A, d0);
        s.pt[0].pb = dist_core_support<BIG>(B, -d0);
        s.lam[0] = 1.0;
        s.n = 1;
    }
    d3 v = P(s.pt[0]);
    double vv = sqnorm(v);
    bool inside = false;
    for (int it = 0; it < DIST_MAX_ITERS; ++it)
    {
        if (!(vv > 0.0)) // the cores touch (or the input is not finite: no distance to report)
        {
            inside = true;
            break;
        }
        SupportPt w;
        w.pa = dist_core_support<BIG>(A, -v);
        w.pb = dist_core_support<BIG>(B, v);
        const d3 wp = P(w);
        const double vw = dot(v, wp);
        // lower and upper bound have met — to 1e-12, or to the rounding error of v·w itself (running error bound of
        // the dot product of v with pa − pb): a support point that is "better" by less than that is noise, and adding
        // it would only flatten the simplex
        const double noise = (fabs(v.x) * (fabs(w.pa.x) + fabs(w.pb.x)) + fabs(v.y) * (fabs(w.pa.y) + fabs(w.pb.y))) +
                             fabs(v.z) * (fabs(w.pa.z) + fabs(w.pb.z));
        if (vv - vw <= DIST_REL_GAP * vv + 1e-15 * noise) break;
        bool seen = false;
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            if (i < s.n)
            {
                const d3 q = P(s.pt[i]);
                seen = seen || (q.x == wp.x && q.y == wp.y && q.z == wp.z);
            }
        }
        if (seen) break; // the support point is a vertex of the simplex already: v is optimal
        DistSimplex t = s;
#pragma unroll
        for (int m = 0; m < 4; ++m)
            if (m == t.n) t.pt[m] = w;
        ++t.n;
        const int rc = dist_reduce(t);
        if (rc == DIST_INSIDE)
        {
            inside = true;
            break;
        }
        if (rc == DIST_STALLED) break;
        const d3 nv = dist_point(t);
        const double nvv = sqnorm(nv);
        if (!(nvv < vv)) break; // rounding: no descent any more, keep the previous point
        s = t;
        v = nv;
        vv = nvv;
    }
    distance = 0.0;
    out_a = out_b = d3{0.0, 0.0, 0.0};
    if (inside) return false;
    const double len = sqrt(vv);
    const double ra = dist_margin(A), rb = dist_margin(B);
    const double dist = (len - ra) - rb;
    if (!(dist > 0.0) || !(dist < 1e300)) return false; // touching / overlapping margins; non-finite input: nothing to report
    d3 a{0.0, 0.0, 0.0}, b{0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        if (i < s.n)
        {
            a = a + s.lam[i] * s.pt[i].pa;
            b = b + s.lam[i] * s.pt[i].pb;
        }
    }
    const d3 n = v * (1.0 / len); // from b towards a
    out_a = (a - ra * n) + O;
    out_b = (b + rb * n) + O;
    distance = dist;
    return true;
}

// 3 blocks of 128 per SM (168 registers, the two simplices partly on the stack) against the 2 that 254 registers allow:
// 7.9 → 7.2 ms over the 14.25 M pairs of C3; 4 blocks (128 registers): 7.3
template <bool BIG>
__global__ void __launch_bounds__(128, 3)
gjk_distance_kernel(BodyArrays bodies, const uint32_t *__restrict__ pair_a, const uint32_t *__restrict__ pair_b, uint64_t n,
                    uint32_t n_bodies, DistanceRec *__restrict__ out, uint8_t *__restrict__ separated)
{
    const uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (k >= n) return;
    const uint32_t ia = pair_a[k], ib = pair_b[k];
    DistanceRec r;
    r.key = (static_cast<uint64_t>(ia) << 32) | ib;
    r.distance = 0.0;
    d3 a{0.0, 0.0, 0.0}, b{0.0, 0.0, 0.0};
    bool sep = false;
    if (ia < n_bodies && ib < n_bodies) // (pk_gjk_distance_batch_device cannot check a list that lives in HBM)
    {
        const ShapeView A = load_shape(bodies, ia);
        const ShapeView B = load_shape(bodies, ib);
        sep = gjk_distance_pair<BIG>(A, B, r.distance, a, b);
    }
    r.point_a[0] = a.x;
    r.point_a[1] = a.y;
    r.point_a[2] = a.z;
    r.point_b[0] = b.x;
    r.point_b[1] = b.y;
    r.point_b[2] = b.z;
    out[k] = r;
    separated[k] = sep ? 1 : 0;
}

} // namespace pk
