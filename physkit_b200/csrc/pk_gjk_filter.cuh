// pk_gjk_filter.cuh — K7a-0: which candidate pairs can gjk_collision (reference src/collision.cpp:165-189) only
// answer with "no"?  Decided in FP32, with a certificate, before any FP64 work is spent on the pair.
//
// Four of five candidate pairs of a broadphase are misses, and a miss has no output other than the missing
// contact.  gjk_collision returns true in two places only: the first support point lies within 1e-6 of the origin
// (collision.cpp:174), or a tetrahedron of support points encloses it (collision.cpp:90-147); both need a point of
// the Minkowski difference A ⊖ B within rounding distance (1e-6, resp. ≈1e-15 relative) of the origin.  If some
// direction d is known with
//        h(d) = max over A ⊖ B of v·d  ≤  −δ‖d‖      for a δ far above those distances,
// every point of A ⊖ B is at least δ from the origin and the reference's answer is "no" whatever path its
// iteration takes.  gjk_filter_kernel looks for such a d in single precision — its decisions need not match the
// reference's, any d will do:
//   * spheres, boxes (OBB, world box, mesh::box hull): one test that is complete for these kinds — the fifteen axes of
//     two boxes and the distance from one centre to the other's box (analytic_separated)
//   * pairs with a general hull: the centre line, then a few steps of a GJK iteration of its own (directions_separate)
// and accepts when the FP32 gap exceeds margin = 1e-4·S + 2e-6, S = ‖cB − cA‖₁ + extent(A) + extent(B) bounding every
// intermediate magnitude: the FP32 evaluation (inputs rounded from FP64 relative to A's centre, a few dozen
// operations, a support vertex chosen in FP32 that may be second best by rounding) is off by less than 1e-5·S, a
// tenth of the margin.  Everything else — hits, near misses, bodies whose quaternion is not unit to 1e-5 — goes to
// gjk_kernel, which runs the reference's iteration in FP64 from its first support.
//
// BASELINE C3 (1 M spheres and boxes): 1,804,818 of 14,245,650 candidate pairs reach gjk_kernel, 1,802,288 of them
// are hits; round 1's FP64 prefilter (the reference's first two supports) let 5,763,139 through.  GJK stage 2.43 →
// 1.24 ms.  tests/test_gjk_filter_host.py runs this file's per-pair decision on the host against the oracle
// (grazing pairs at gaps from 1e-12 of their size, all kinds, coordinates up to 1e6).
#pragma once

#include "pk_narrowphase.cuh"

namespace pk
{

struct f3
{
    float x, y, z;
};
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ float dotf(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 crossf(f3 a, f3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ float norm1f(f3 a) { return fabsf(a.x) + fabsf(a.y) + fabsf(a.z); }
__device__ __forceinline__ f3 absf3(f3 a) { return {fabsf(a.x), fabsf(a.y), fabsf(a.z)}; }
__device__ __forceinline__ f3 to_f3(d3 a) { return {static_cast<float>(a.x), static_cast<float>(a.y), static_cast<float>(a.z)}; }
// unit vector, or zero when the input is zero / not finite (the caller gives the pair up)
__device__ __forceinline__ f3 unitf(f3 a)
{
    const float z = dotf(a, a);
    if (!(z > 1e-30f && z < 1e30f)) return {0.f, 0.f, 0.f};
    return rsqrtf(z) * a;
}

// One body for the filter: centre relative to a reference point near the pair, a box / sphere / vertex cloud around it.
struct ShapeF
{
    int kind;        // KIND_OBB (also world boxes and mesh::box hulls), KIND_SPHERE, KIND_HULL
    uint32_t nverts; // HULL
    const float4 *vf;
    f3 c, h;         // centre; OBB: half extents
    float rho;       // SPHERE: radius (its h is 0); 0 for the others
    float qx, qy, qz, qw;
    float ext;       // bound of ‖x − c‖ over the shape
    bool ok;         // false: do not filter pairs with this body
};

__device__ __forceinline__ d3 shape_centre(const ShapeView &s)
{
    if (s.kind == KIND_AABB) return d3{0.5 * (s.p.x + s.h.x), 0.5 * (s.p.y + s.h.y), 0.5 * (s.p.z + s.h.z)};
    return s.p;
}

__device__ __forceinline__ ShapeF shape_f(const ShapeView &s, d3 origin)
{
    ShapeF f;
    f.kind = s.kind;
    f.nverts = s.nverts;
    f.vf = s.vf;
    f.rho = 0.f;
    f.c = to_f3(shape_centre(s) - origin);
    f.qx = static_cast<float>(s.q.x);
    f.qy = static_cast<float>(s.q.y);
    f.qz = static_cast<float>(s.q.z);
    f.qw = static_cast<float>(s.q.w);
    const float qq = f.qx * f.qx + f.qy * f.qy + f.qz * f.qz + f.qw * f.qw;
    f.ok = fabsf(qq - 1.f) < 1e-5f; // (the box tests assume an orthonormal frame; off by more: not filtered)
    if (s.kind == KIND_AABB)
    {
        f.kind = KIND_OBB; // (identity quaternion from load_shape)
        f.h = absf3(to_f3(d3{0.5 * (s.h.x - s.p.x), 0.5 * (s.h.y - s.p.y), 0.5 * (s.h.z - s.p.z)}));
        f.ext = norm1f(f.h);
    }
    else if (s.kind == KIND_OBB)
    {
        f.h = absf3(to_f3(s.h)); // (support() picks among the corners (±hx, ±hy, ±hz) whatever the signs stored)
        f.ext = norm1f(f.h);
    }
    else if (s.kind == KIND_SPHERE)
    {
        f.h = {0.f, 0.f, 0.f};
        f.rho = fabsf(static_cast<float>(s.h.x)); // (the reference's support is p + r·d̂ whatever the sign of r; a negative
                                                  // radius makes the ball of radius |r| all the same)
        f.ext = f.rho;
        f.qx = f.qy = f.qz = 0.f; // the frame of a ball is anybody's
        f.qw = 1.f;
        f.ok = true;
    }
    else if (s.hull_r < 0.f)
    {
        f.kind = KIND_OBB; // mesh::box: vertices (±hx, ±hy, ±hz), s.h holds the local minimum
        f.h = absf3(to_f3(s.h));
        f.ext = norm1f(f.h);
    }
    else
    {
        f.h = {0.f, 0.f, 0.f};
        f.ext = 1.75f * s.hull_r; // hull_r = largest |coordinate|, rounded up: ‖v‖₂ ≤ √3·hull_r
    }
    if (!(f.ext < 1e18f) || !(norm1f(f.c) < 1e18f)) f.ok = false;
    return f;
}

// Eigen's q·v·q⁻¹ (lin_alg.h:493-499) in single precision, same operation order as rotate() in pk_common.cuh
__device__ __forceinline__ f3 rotatef(float qx, float qy, float qz, float qw, f3 v)
{
    const f3 qv{qx, qy, qz};
    f3 uv = crossf(qv, v);
    uv = uv + uv;
    const f3 c = crossf(qv, uv);
    return {(v.x + qw * uv.x) + c.x, (v.y + qw * uv.y) + c.y, (v.z + qw * uv.z) + c.z};
}

// a point of the shape that is farthest along the unit vector d, to FP32 accuracy
__device__ __forceinline__ f3 supportf(const ShapeF &s, f3 d)
{
    if (s.kind == KIND_SPHERE) return s.c + s.rho * d;
    const f3 l = rotatef(-s.qx, -s.qy, -s.qz, s.qw, d);
    f3 v;
    if (s.kind == KIND_OBB)
        v = {l.x >= 0.f ? s.h.x : -s.h.x, l.y >= 0.f ? s.h.y : -s.h.y, l.z >= 0.f ? s.h.z : -s.h.z};
    else
    {
        float best = -3.4e38f;
        v = {0.f, 0.f, 0.f};
        const float4 *__restrict__ vf = s.vf;
        for (uint32_t i = 0; i < s.nverts; ++i)
        {
            const float4 w = __ldg(vf + i);
            const float t = w.x * l.x + w.y * l.y + w.z * l.z;
            if (t > best)
            {
                best = t;
                v = {w.x, w.y, w.z};
            }
        }
    }
    return s.c + rotatef(s.qx, s.qy, s.qz, s.qw, v);
}

__device__ __forceinline__ f3 minkowskif(const ShapeF &a, const ShapeF &b, f3 d) { return supportf(a, d) - supportf(b, -d); }

// Next search direction from the simplex p[0..n) (newest last), the regions of collision.cpp:12-147 in single
// precision.  Returns false when the origin looks enclosed or the direction degenerates: the pair is not filtered.
__device__ __forceinline__ bool simplex_dir_f(f3 *p, int &n, f3 &dir)
{
    if (n == 4)
    {
        const f3 a = p[3], b = p[2], c = p[1], d = p[0];
        const f3 ao = -a;
        f3 abc = crossf(b - a, c - a), acd = crossf(c - a, d - a), adb = crossf(d - a, b - a);
        if (dotf(abc, d - a) > 0.f) abc = -abc;
        if (dotf(acd, b - a) > 0.f) acd = -acd;
        if (dotf(adb, c - a) > 0.f) adb = -adb;
        if (dotf(abc, ao) > 0.f)
        {
            p[0] = c;
            p[1] = b;
        }
        else if (dotf(acd, ao) > 0.f)
        {
            p[0] = d;
            p[1] = c;
        }
        else if (dotf(adb, ao) > 0.f)
        {
            p[0] = b;
            p[1] = d;
        }
        else
            return false;
        p[2] = a;
        n = 3;
    }
    f3 raw;
    if (n == 3)
    {
        const f3 a = p[2], b = p[1], c = p[0];
        const f3 ab = b - a, ac = c - a, ao = -a;
        const f3 abc = crossf(ab, ac);
        if (dotf(crossf(ab, abc), ao) > 0.f)
        {
            p[0] = b; // {b, a}
            p[1] = a;
            n = 2;
            raw = crossf(crossf(ab, ao), ab);
        }
        else if (dotf(crossf(abc, ac), ao) > 0.f)
        {
            p[1] = a; // {c, a}
            n = 2;
            raw = crossf(crossf(ac, ao), ac);
        }
        else if (dotf(abc, ao) <= 0.f)
        {
            p[0] = b;
            p[1] = c;
            raw = -abc;
        }
        else
            raw = abc;
    }
    else
    {
        const f3 a = p[1], b = p[0];
        const f3 ab = b - a, ao = -a;
        if (dotf(ab, ao) > 0.f)
            raw = crossf(crossf(ab, ao), ab);
        else
        {
            p[0] = a;
            n = 1;
            raw = ao;
        }
    }
    dir = unitf(raw);
    return dotf(dir, dir) > 0.5f;
}

#ifndef PK_GJK_FILTER_ITERS
#define PK_GJK_FILTER_ITERS 2 // directions after the first two, pairs with a general hull
#endif

// ---- one complete test for the analytic kinds ----
// Every analytic shape is a box widened by a ball: OBB / world box / mesh::box hull = (half extents h, radius 0), sphere =
// (h = 0, radius r).  In A's frame (R = B's axes there, t = B's centre) three families of certificates, all evaluated for
// every pair so that the lanes of a warp stay together whatever the kinds are:
//   * the fifteen axes of two oriented boxes: an axis L (‖L‖ ≤ 1) with |t·L| − r_A(L) − r_B(L) − (ρ_A + ρ_B) > margin
//     separates the shapes by more than the margin (complete for box–box)
//   * distance from B's centre to A's box, minus B's bounding radius and A's ball (exact when B is a sphere)
//   * the same with the roles swapped; B's centre in A's frame and A's centre in B's frame fall out of the axis tests
__device__ __forceinline__ bool analytic_separated(const ShapeF &a, const ShapeF &b, float margin)
{
    // q = conj(qa) ⊗ qb, t = conj(qa) (cb − ca) qa
    const float ax = -a.qx, ay = -a.qy, az = -a.qz, aw = a.qw;
    const float x = aw * b.qx + b.qw * ax + (ay * b.qz - az * b.qy);
    const float y = aw * b.qy + b.qw * ay + (az * b.qx - ax * b.qz);
    const float z = aw * b.qz + b.qw * az + (ax * b.qy - ay * b.qx);
    const float w = aw * b.qw - (ax * b.qx + ay * b.qy + az * b.qz);
    const f3 t = rotatef(ax, ay, az, aw, b.c - a.c);
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
    const float R00 = 1.f - 2.f * (yy + zz), R01 = 2.f * (xy - wz), R02 = 2.f * (xz + wy);
    const float R10 = 2.f * (xy + wz), R11 = 1.f - 2.f * (xx + zz), R12 = 2.f * (yz - wx);
    const float R20 = 2.f * (xz - wy), R21 = 2.f * (yz + wx), R22 = 1.f - 2.f * (xx + yy);
    const float A00 = fabsf(R00), A01 = fabsf(R01), A02 = fabsf(R02), A10 = fabsf(R10), A11 = fabsf(R11), A12 = fabsf(R12),
                A20 = fabsf(R20), A21 = fabsf(R21), A22 = fabsf(R22);
    const float a0 = a.h.x, a1 = a.h.y, a2 = a.h.z, b0 = b.h.x, b1 = b.h.y, b2 = b.h.z;
    const float balls = a.rho + b.rho;
    // B's centre along A's axes, A's centre along B's (up to sign)
    const float ta0 = fabsf(t.x), ta1 = fabsf(t.y), ta2 = fabsf(t.z);
    const float tb0 = fabsf(fmaf(t.x, R00, fmaf(t.y, R10, t.z * R20))), tb1 = fabsf(fmaf(t.x, R01, fmaf(t.y, R11, t.z * R21))),
                tb2 = fabsf(fmaf(t.x, R02, fmaf(t.y, R12, t.z * R22)));
    float g; // largest gap over the box axes
    g = ta0 - (a0 + fmaf(b0, A00, fmaf(b1, A01, b2 * A02)));
    g = fmaxf(g, ta1 - (a1 + fmaf(b0, A10, fmaf(b1, A11, b2 * A12))));
    g = fmaxf(g, ta2 - (a2 + fmaf(b0, A20, fmaf(b1, A21, b2 * A22))));
    g = fmaxf(g, tb0 - (b0 + fmaf(a0, A00, fmaf(a1, A10, a2 * A20))));
    g = fmaxf(g, tb1 - (b1 + fmaf(a0, A01, fmaf(a1, A11, a2 * A21))));
    g = fmaxf(g, tb2 - (b2 + fmaf(a0, A02, fmaf(a1, A12, a2 * A22))));
    // A_i × B_j
    g = fmaxf(g, fabsf(t.z * R10 - t.y * R20) - (fmaf(a1, A20, a2 * A10) + fmaf(b1, A02, b2 * A01)));
    g = fmaxf(g, fabsf(t.z * R11 - t.y * R21) - (fmaf(a1, A21, a2 * A11) + fmaf(b0, A02, b2 * A00)));
    g = fmaxf(g, fabsf(t.z * R12 - t.y * R22) - (fmaf(a1, A22, a2 * A12) + fmaf(b0, A01, b1 * A00)));
    g = fmaxf(g, fabsf(t.x * R20 - t.z * R00) - (fmaf(a0, A20, a2 * A00) + fmaf(b1, A12, b2 * A11)));
    g = fmaxf(g, fabsf(t.x * R21 - t.z * R01) - (fmaf(a0, A21, a2 * A01) + fmaf(b0, A12, b2 * A10)));
    g = fmaxf(g, fabsf(t.x * R22 - t.z * R02) - (fmaf(a0, A22, a2 * A02) + fmaf(b0, A11, b1 * A10)));
    g = fmaxf(g, fabsf(t.y * R00 - t.x * R10) - (fmaf(a0, A10, a1 * A00) + fmaf(b1, A22, b2 * A21)));
    g = fmaxf(g, fabsf(t.y * R01 - t.x * R11) - (fmaf(a0, A11, a1 * A01) + fmaf(b0, A22, b2 * A20)));
    g = fmaxf(g, fabsf(t.y * R02 - t.x * R12) - (fmaf(a0, A12, a1 * A02) + fmaf(b0, A21, b1 * A20)));
    // centre of one to the box of the other
    const float ex = fmaxf(ta0 - a0, 0.f), ey = fmaxf(ta1 - a1, 0.f), ez = fmaxf(ta2 - a2, 0.f);
    const float fx = fmaxf(tb0 - b0, 0.f), fy = fmaxf(tb1 - b1, 0.f), fz = fmaxf(tb2 - b2, 0.f);
    const float reach_b = sqrtf(fmaf(b0, b0, fmaf(b1, b1, b2 * b2))) * 1.000001f + balls + margin; // B inside this ball around its centre
    const float reach_a = sqrtf(fmaf(a0, a0, fmaf(a1, a1, a2 * a2))) * 1.000001f + balls + margin;
    return g - balls > margin || fmaf(ex, ex, fmaf(ey, ey, ez * ez)) > reach_b * reach_b || fmaf(fx, fx, fmaf(fy, fy, fz * fz)) > reach_a * reach_a;
}

// any two shapes: directions tried one after the other — centre to centre, then a GJK iteration of its own
__device__ __forceinline__ bool directions_separate(const ShapeF &a, const ShapeF &b, float margin, int iters)
{
    f3 p[4];
    int n = 1;
    f3 dir = unitf(b.c - a.c);
    if (!(dotf(dir, dir) > 0.5f)) dir = {1.f, 0.f, 0.f};
    p[0] = minkowskif(a, b, dir);
    if (dotf(p[0], dir) < -margin) return true;
    dir = unitf(-p[0]);
    bool separated = false, go = dotf(dir, dir) > 0.5f;
    for (int iter = 0; go && iter <= iters; ++iter)
    {
        const f3 q = minkowskif(a, b, dir);
        const float t = dotf(q, dir);
        if (t < -margin)
        {
            separated = true;
            go = false;
        }
        else if (!(t > margin))
            go = false; // touching to within the margin: FP64 decides
        else
        {
            p[n++] = q;
            go = simplex_dir_f(p, n, dir);
        }
    }
    return separated;
}

// true: the pair is separated for certain (see the head of this file)
__device__ __forceinline__ bool certainly_separated(const ShapeView &A, const ShapeView &B, int iters)
{
    const d3 origin = shape_centre(A);
    const ShapeF a = shape_f(A, origin), b = shape_f(B, origin);
    if (!(a.ok && b.ok)) return false;
    const float S = norm1f(b.c) + norm1f(a.c) + a.ext + b.ext;
    const float margin = 1e-4f * S + 2e-6f;
    if (a.kind != KIND_HULL && b.kind != KIND_HULL) return analytic_separated(a, b, margin);
    return directions_separate(a, b, margin, iters);
}

// The FP64 form of the stage's first kernel, for contexts that hold many-vertex hulls (and, with PK_GJK_EXACT_PREFILTER=1,
// for A/B runs against round 1): the first two support evaluations of gjk_collision for every pair
// (collision.cpp:170-182).  A pair whose second support makes no progress is separated; it gets hit = 0 and never
// reaches the divergent part.  All lanes do identical work; survivors are appended to a work list per shape-kind
// class.  A hull's support is a scan of its vertices in FP32 followed by an exact decision (pk_common.cuh), so a filter
// in FP32 costs what this costs and its survivors would pay for their first two supports a second time (C4: 6.7 against
// 6.0 ms) — instead those two points are carried to gjk_kernel (CARRY, 96 bytes per survivor).  FILTER: pairs of
// analytic shapes in such a context are still settled by analytic_separated first.
template <bool CARRY, bool FILTER = false>
__global__ void __launch_bounds__(128)
gjk_prefilter_kernel(BodyArrays bodies, const uint64_t *__restrict__ keys, const uint32_t *__restrict__ pair_a,
                     const uint32_t *__restrict__ pair_b, uint64_t npairs_cap, const unsigned long long *__restrict__ npairs_dev,
                     uint8_t *__restrict__ hit, uint32_t *__restrict__ work /*[4][work_stride]*/, uint64_t work_stride,
                     unsigned long long *__restrict__ work_count /*[4]*/, GjkCarry *__restrict__ carry /*[npairs]*/)
{
    const uint64_t npairs = device_count(npairs_dev, npairs_cap);
    const uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    bool survive = false;
    uint32_t cls = 0;
    if (k < npairs)
    {
        uint32_t ia, ib;
        load_pair(keys, pair_a, pair_b, k, ia, ib);
        ShapeView A = load_shape(bodies, ia);
        ShapeView B = load_shape(bodies, ib);
        cls = (A.kind == KIND_SPHERE ? 1u : 0u) | (B.kind == KIND_SPHERE ? 2u : 0u);
        const bool filtered = FILTER && A.kind != KIND_HULL && B.kind != KIND_HULL && certainly_separated(A, B, 0);
        SupportPt s0;
        d3 p0{1.0, 0.0, 0.0};
        if (!filtered)
        {
            s0 = minkowski_support<CARRY>(A, B, d3{1.0, 0.0, 0.0});
            p0 = P(s0);
        }
        double2 *c = reinterpret_cast<double2 *>(carry + k);
        if (filtered)
            survive = false;
        else if (sqnorm(p0) < 1e-12)
        {
            survive = true; // origin hit on the first point: a hit with a one-point simplex (collision.cpp:174)
            if constexpr (CARRY)
            {
                c[0] = make_double2(s0.pa.x, s0.pa.y);
                c[1] = make_double2(s0.pa.z, s0.pb.x);
                c[2] = make_double2(s0.pb.y, s0.pb.z);
                c[3] = make_double2(__longlong_as_double(0x7FF8000000000000ll), 0.0);
            }
        }
        else
        {
            const d3 dir = -normalized(p0);
            const SupportPt s1 = minkowski_support<CARRY>(A, B, dir);
            survive = !(dot(P(s1), dir) <= 0.0); // collision.cpp:181-182
            if (CARRY && survive)
            {
                c[0] = make_double2(s0.pa.x, s0.pa.y);
                c[1] = make_double2(s0.pa.z, s0.pb.x);
                c[2] = make_double2(s0.pb.y, s0.pb.z);
                c[3] = make_double2(s1.pa.x, s1.pa.y);
                c[4] = make_double2(s1.pa.z, s1.pb.x);
                c[5] = make_double2(s1.pb.y, s1.pb.z);
            }
        }
        if (!survive) hit[k] = 0;
    }
    // one atomic per (warp, class); dead lanes take a class of their own
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, survive ? cls : 4u);
    if (survive)
    {
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(peers) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(work_count + cls, static_cast<unsigned long long>(__popc(peers)));
        base = __shfl_sync(peers, base, leader);
        work[cls * work_stride + base + __popc(peers & ((1u << lane) - 1u))] = static_cast<uint32_t>(k);
    }
}

// One thread per candidate pair; survivors are listed per shape-kind class (bit 0: A is a sphere, bit 1: B is a sphere)
// so that the lanes of a gjk_kernel warp run the same support code.
// 8 blocks per SM (64 registers, 56 bytes of stack) instead of the 5 that 88 registers allow: the kernel is bound by its
// FP32 instruction stream and the gathers it waits for, more warps cover both (C3 GJK stage 1.28 → 1.19 ms; 6 blocks: 1.23)
#ifndef PK_GJK_FILTER_MIN_BLOCKS
#define PK_GJK_FILTER_MIN_BLOCKS 8
#endif
__global__ void __launch_bounds__(128, PK_GJK_FILTER_MIN_BLOCKS)
gjk_filter_kernel(BodyArrays bodies, const uint64_t *__restrict__ keys, const uint32_t *__restrict__ pair_a,
                  const uint32_t *__restrict__ pair_b, uint64_t npairs_cap, const unsigned long long *__restrict__ npairs_dev,
                  uint8_t *__restrict__ hit, uint32_t *__restrict__ work /*[4][work_stride]*/, uint64_t work_stride,
                  unsigned long long *__restrict__ work_count /*[4]*/, int iters)
{
    const uint64_t npairs = device_count(npairs_dev, npairs_cap);
    const uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    bool survive = false;
    uint32_t cls = 0;
    if (k < npairs)
    {
        uint32_t ia, ib;
        load_pair(keys, pair_a, pair_b, k, ia, ib);
        const ShapeView A = load_shape(bodies, ia);
        const ShapeView B = load_shape(bodies, ib);
        cls = (A.kind == KIND_SPHERE ? 1u : 0u) | (B.kind == KIND_SPHERE ? 2u : 0u);
        survive = !certainly_separated(A, B, iters);
        if (!survive) hit[k] = 0;
    }
    // one atomic per (warp, class); dead lanes take a class of their own
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, survive ? cls : 4u);
    if (survive)
    {
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(peers) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(work_count + cls, static_cast<unsigned long long>(__popc(peers)));
        base = __shfl_sync(peers, base, leader);
        work[cls * work_stride + base + __popc(peers & ((1u << lane) - 1u))] = static_cast<uint32_t>(k);
    }
}

} // namespace pk
