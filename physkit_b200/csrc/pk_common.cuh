// pk_common.cuh — device-side algebra, shape records and support functions.
//
// Every translation unit of this library is compiled with -fmad=false: the reference is built
// with plain -O3 for x86-64 (CMakeLists.txt:126-132 → SSE2, no FMA), so a*b+c must round twice
// for results to be bit-identical.  Operation order follows Eigen's fixed-size double kernels as
// used by include/physkit/algebra/lin_alg.h (citations relative to /root/reference).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace pk
{

constexpr int KIND_AABB = 0;   // aabb::support            bounds.h:164-174 (pose-less)
constexpr int KIND_OBB = 1;    // obb::support             bounds.h:539-548
constexpr int KIND_SPHERE = 2; // bounding_sphere::support bounds.h:328-329
constexpr int KIND_HULL = 3;   // mesh::instance::support  src/mesh.cpp:442-448, 341-358

constexpr uint32_t HULL_PREFILTER_MIN = 16; // hulls up to this size are scanned in FP64 directly

#ifndef PK_HULL_UNROLL
#define PK_HULL_UNROLL 4
#endif
constexpr int HULL_UNROLL = PK_HULL_UNROLL; // 32-byte vertex-pair loads in flight per lane in the float scan of a hull

constexpr uint8_t FLAG_STATIC = 1;
constexpr uint8_t FLAG_ALIVE = 2;

struct d3
{
    double x, y, z;
};

#define PK_HD __host__ __device__ __forceinline__

PK_HD d3 make_d3(double x, double y, double z) { return d3{x, y, z}; }
PK_HD d3 operator+(d3 a, d3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
PK_HD d3 operator-(d3 a, d3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
PK_HD d3 operator-(d3 a) { return {-a.x, -a.y, -a.z}; }
PK_HD d3 operator*(d3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
PK_HD d3 operator*(double s, d3 a) { return {s * a.x, s * a.y, s * a.z}; }
// Eigen 3-vector redux with SSE2 packets: (x0 + x1) + x2   (lin_alg.h:212-213, :229)
PK_HD double dot(d3 a, d3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
PK_HD double sqnorm(d3 a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }
// Eigen cross3 (lin_alg.h:215-219)
PK_HD d3 cross(d3 a, d3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
// a / s for |a|, s in a safe exponent range, bit-identical to IEEE division (round to nearest), given
// r = RN(1/s).  Markstein's division: q0 = RN(a·r) is within 1.5 ulp, one residual step makes it faithful,
// and for a faithful q1 the residual a − s·q1 is exact and RN(q1 + residual·r) is the correctly rounded
// quotient because r is the correctly rounded reciprocal (P. Markstein, IBM J. R&D 34(1), 1990, Thm 8;
// the same scheme finishes div.rn.f64).  Five instructions per quotient instead of ≈26, and the three
// components of a vector share the reciprocal.  tests/test_gpu_division.py hammers it against `/`.
__device__ __forceinline__ double pk_div_by_rcp(double a, double s, double r)
{
    const double q0 = __dmul_rn(a, r);
    const double e0 = __fma_rn(-s, q0, a);
    const double q1 = __fma_rn(e0, r, q0);
    const double e1 = __fma_rn(-s, q1, a);
    return __fma_rn(e1, r, q1);
}

// Eigen normalized(): z > 0 ? v / sqrt(z) : v, true division (lin_alg.h:232-240)
PK_HD d3 normalized(d3 a)
{
    double z = sqnorm(a);
    if (z > 0.0)
    {
        double s = sqrt(z);
#ifdef __CUDA_ARCH__
#ifndef PK_TRUE_DIVISION
        // Fast path: s and every non-zero component far from overflow / underflow, so that neither the
        // quotients nor the residuals leave the normal range.  z in [1e-200, 1e200] bounds s and the
        // components from above; a zero component divides to itself (s is positive and finite).
        const double ax = fabs(a.x), ay = fabs(a.y), az = fabs(a.z);
        const bool safe = z > 1e-200 && z < 1e200 && (ax > 1e-150 || ax == 0.0) && (ay > 1e-150 || ay == 0.0) &&
                          (az > 1e-150 || az == 0.0);
        if (safe)
        {
            const double r = __drcp_rn(s);
            return {ax == 0.0 ? a.x : pk_div_by_rcp(a.x, s, r), ay == 0.0 ? a.y : pk_div_by_rcp(a.y, s, r),
                    az == 0.0 ? a.z : pk_div_by_rcp(a.z, s, r)};
        }
#endif
#endif
        return {a.x / s, a.y / s, a.z / s};
    }
    return a;
}
PK_HD double dmin(double a, double b) { return (b < a) ? b : a; } // std::min
PK_HD double dmax(double a, double b) { return (a < b) ? b : a; } // std::max

struct dq
{
    double x, y, z, w; // Eigen coeffs order (lin_alg.h:388)
};
PK_HD dq conjugate(dq q) { return {-q.x, -q.y, -q.z, q.w}; }
// Eigen _transformVector (lin_alg.h:493-499): uv = qv×v; uv += uv; v + w·uv + qv×uv (left-assoc)
PK_HD d3 rotate(dq q, d3 v)
{
    d3 qv{q.x, q.y, q.z};
    d3 uv = cross(qv, v);
    uv = uv + uv;
    d3 c = cross(qv, uv);
    return {(v.x + q.w * uv.x) + c.x, (v.y + q.w * uv.y) + c.y, (v.z + q.w * uv.z) + c.z};
}

// 64-byte device shape record.
//   AABB  : a = min,  b = max                       (local box = a,b)
//   OBB   : a = half                                (local box = ±a)
//   SPHERE: a.x = r                                 (local box = ±r)
//   HULL  : a = local min, b = local max, vertices  (aabb::from_points, mesh.h:191)
struct alignas(16) ShapeRec
{
    double a[3];
    double b[3];
    int32_t kind;
    uint32_t vert_off;
    uint32_t nverts;
    uint32_t mesh_box; // HULL: 1 when the vertices are exactly mesh::box's table (src/mesh.cpp:31-40) for half = b
};
static_assert(sizeof(ShapeRec) == 64, "ShapeRec must be one 64-byte line");

// What a support query needs for one body, in registers.
struct ShapeView
{
    int kind;
    uint32_t nverts;
    const double *verts; // HULL: xyz triples (3 doubles per vertex)
    const float4 *vf;    // HULL: the same vertices rounded to float (x,y,z,0) for the argmax prefilter
    d3 p;                // AABB: min | others: position / centre
    d3 h;                // AABB: max | OBB: half | SPHERE: h.x = r
    float hull_r;        // HULL: largest |coordinate| of the local box (error bound of the prefilter); < 0: mesh::box
    dq q;                // OBB / HULL
};

struct BodyArrays
{
    const ShapeRec *shapes;
    const double *verts;
    const float4 *verts_f;
    const double *pos;
    const double *quat;
    const uint32_t *shape_id;
};

__device__ __forceinline__ ShapeView load_shape(const BodyArrays &ba, uint32_t body)
{
    const ShapeRec *__restrict__ shapes = ba.shapes;
    const double *__restrict__ verts = ba.verts;
    const double *__restrict__ pos = ba.pos;
    const double *__restrict__ quat = ba.quat;
    const uint32_t *__restrict__ shape_id = ba.shape_id;
    ShapeView v;
    const ShapeRec *s = shapes + shape_id[body];
    // 64-byte record as four 16-byte loads
    const double2 *sp = reinterpret_cast<const double2 *>(s);
    double2 s0 = __ldg(sp), s1 = __ldg(sp + 1), s2 = __ldg(sp + 2);
    int4 s3 = __ldg(reinterpret_cast<const int4 *>(sp + 3));
    v.kind = s3.x;
    v.nverts = static_cast<uint32_t>(s3.z);
    v.verts = verts + 3ull * static_cast<uint32_t>(s3.y);
    v.vf = ba.verts_f + static_cast<uint32_t>(s3.y);
    v.hull_r = 0.f;
    if (v.kind == KIND_HULL)
    {
        double r = fmax(fmax(fmax(fabs(s0.x), fabs(s0.y)), fmax(fabs(s1.x), fabs(s1.y))), fmax(fabs(s2.x), fabs(s2.y)));
        v.hull_r = s3.w == 1 ? -1.f : __double2float_ru(r);
    }
    if (v.kind == KIND_AABB)
    {
        v.p = {s0.x, s0.y, s1.x};
        v.h = {s1.y, s2.x, s2.y};
        v.q = {0, 0, 0, 1};
    }
    else
    {
        v.p = {pos[3ull * body], pos[3ull * body + 1], pos[3ull * body + 2]};
        v.h = {s0.x, s0.y, s1.x};
        const double2 *qp = reinterpret_cast<const double2 *>(quat + 4ull * body);
        double2 q0 = __ldg(qp), q1 = __ldg(qp + 1);
        v.q = {q0.x, q0.y, q1.x, q1.y};
    }
    return v;
}

// ---- support of a many-vertex hull: float prefilter, exact decision ----
// F_i = float dot of the float-rounded vertex with the float-rounded direction differs from the reference's
// FP64 dot D_i by at most
//   E = 5·2⁻²⁴ · (|x·lx| + |y·ly| + |z·lz|)  ≤  1e-6 · R · ‖l‖₁      (R = hull's largest |coordinate|)
// (two input roundings + three float operations, 1e-6 leaves a 3× margin).  Any vertex with
// F_i < max F − 2E has D_i < D_argmaxF, so it can be neither the maximum nor tied with it.  One pass keeps
// the largest and the second largest F: when the runner-up is below the threshold the float argmax is the
// only candidate and IS the reference's answer (the usual case: neighbouring vertices of a 32–256 vertex
// hull differ by ≈1e-2, the threshold is ≈1e-6).  Otherwise a second pass compares the candidates with the
// reference's exact expression, lowest index first (strict '>', src/mesh.cpp:341-358).
// Measured dead end (profiles/r1_c4_hull_support_ab.json): letting the lanes that arrive here together
// scan one hull at a time with consecutive 16-byte loads and a redux.sync maximum for EVERY query is bit-exact
// too but slower (C4, 2 M pairs: 109.5 ms vs 76.3 ms) — the first pass is bound by issued instructions, not by L1
// lines, and serving the group's queries one after the other adds a fixed cost per query.  Only the second pass
// (below) is shared.
__device__ __forceinline__ float pk_hull_fdot(float4 w, float lx, float ly, float lz) { return fmaf(w.x, lx, fmaf(w.y, ly, w.z * lz)); }
__device__ __forceinline__ double pk_shfl_f64(unsigned mask, double x, int src)
{
    unsigned long long b;
    memcpy(&b, &x, 8);
    const unsigned lo = __shfl_sync(mask, static_cast<unsigned>(b), src), hi = __shfl_sync(mask, static_cast<unsigned>(b >> 32), src);
    b = (static_cast<unsigned long long>(hi) << 32) | lo;
    memcpy(&x, &b, 8);
    return x;
}
__device__ __forceinline__ unsigned long long pk_shfl_u64(unsigned mask, unsigned long long x, int src)
{
    const unsigned lo = __shfl_sync(mask, static_cast<unsigned>(x), src), hi = __shfl_sync(mask, static_cast<unsigned>(x >> 32), src);
    return (static_cast<unsigned long long>(hi) << 32) | lo;
}

// The exact comparison among the vertices within 2E of the float maximum, for the lanes `todo` of the coalesced group
// `group` (see hull_argmax), served one after the other by the whole group.  Out of line: it must not cost the
// callers registers (sphere / box scenes never get here).
__device__ __noinline__ uint32_t hull_exact_pass(const float4 *__restrict__ vf, const double *__restrict__ v, uint32_t nverts, float lx, float ly,
                                                 float lz, float thr_own, double ldx, double ldy, double ldz, unsigned group, unsigned todo,
                                                 uint32_t best)
{
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t g = static_cast<uint32_t>(__popc(group));
    const uint32_t r = static_cast<uint32_t>(__popc(group & ((1u << lane) - 1u)));
    uint32_t result = best;
    for (; todo; todo &= todo - 1u)
    {
        const int src = __ffs(static_cast<int>(todo)) - 1;
        const float4 *__restrict__ q = reinterpret_cast<const float4 *>(pk_shfl_u64(group, reinterpret_cast<unsigned long long>(vf), src));
        const double *__restrict__ qv = reinterpret_cast<const double *>(pk_shfl_u64(group, reinterpret_cast<unsigned long long>(v), src));
        const uint32_t n = __shfl_sync(group, nverts, src);
        const float qx = __shfl_sync(group, lx, src), qy = __shfl_sync(group, ly, src), qz = __shfl_sync(group, lz, src);
        const float thr = __shfl_sync(group, thr_own, src);
        const double dx = pk_shfl_f64(group, ldx, src), dy = pk_shfl_f64(group, ldy, src), dz = pk_shfl_f64(group, ldz, src);
        double my_dot = 0.0;
        uint32_t my_i = 0;
        bool have = false;
        for (uint32_t i = r; i < n; i += g) // ascending: the first of equal dots this lane meets has the lowest index
        {
            if (pk_hull_fdot(__ldg(q + i), qx, qy, qz) >= thr)
            {
                const double t = (qv[3 * i] * dx + qv[3 * i + 1] * dy) + qv[3 * i + 2] * dz;
                if (!have || t > my_dot)
                {
                    my_dot = t;
                    my_i = i;
                    have = true;
                }
            }
        }
        unsigned cand = __ballot_sync(group, have);
        double gd = 0.0;
        uint32_t gi = 0;
        bool ghave = false;
        for (; cand; cand &= cand - 1u)
        {
            const int c = __ffs(static_cast<int>(cand)) - 1;
            const double t = pk_shfl_f64(group, my_dot, c);
            const uint32_t i = __shfl_sync(group, my_i, c);
            if (!ghave || t > gd || (t == gd && i < gi))
            {
                gd = t;
                gi = i;
                ghave = true;
            }
        }
        if (static_cast<int>(lane) == src) result = gi;
    }
    return result;
}

__device__ __forceinline__ uint32_t hull_argmax(const float4 *__restrict__ vf, const double *__restrict__ v, uint32_t nverts, float hull_r, d3 l)
{
    const float lx = static_cast<float>(l.x), ly = static_cast<float>(l.y), lz = static_cast<float>(l.z);
    const float E = 1e-6f * hull_r * (fabsf(lx) + fabsf(ly) + fabsf(lz)) + 1e-37f;
    uint32_t best = 0;
    float f1 = -3.4e38f, f2 = -3.4e38f; // largest, second largest (equal values count twice)
    // ncu (C4): the scan saturates the L1 data stage (72–89 % of its peak) — every lane walks its own hull, so
    // a 16-byte load is one wavefront per lane.  Two vertices per 32-byte load (LDG.256, sm_100) halve the
    // wavefronts; hulls start at even vertex offsets (pk_shape_hull) so that the pairs are aligned.
    auto track = [&](const float4 w, uint32_t i)
    {
        const float f = pk_hull_fdot(w, lx, ly, lz);
        const bool gt = f > f1;
        f2 = gt ? f1 : fmaxf(f2, f);
        best = gt ? i : best;
        f1 = gt ? f : f1;
    };
    const uint32_t npairs = nverts >> 1;
#pragma unroll HULL_UNROLL
    for (uint32_t k = 0; k < npairs; ++k)
    {
        float4 a, b;
#ifdef __CUDA_ARCH__
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                     : "l"(vf + 2 * k));
#else // host build of the kernels (tests/cpp/simt_host.h): same two vertices, plain loads
        a = vf[2 * k];
        b = vf[2 * k + 1];
#endif
        track(a, 2 * k);
        track(b, 2 * k + 1);
    }
    if (nverts & 1u) track(__ldg(vf + nverts - 1), nverts - 1);
    // Several candidates within 2E of the maximum: the exact comparison.  This is not the rare case it looks like:
    // EPA asks along face normals, which are perpendicular to edges between hull vertices, so the vertices behind
    // the current face tie to within rounding in most queries.  One lane scanning its whole hull again while the
    // rest of the warp waits was half of the support instructions of BASELINE C4 (ncu: 1.3 of 32 lanes on these
    // lines); instead the lanes that arrive here together serve the ambiguous queries among them one after the
    // other — the query is broadcast, lane r of g scans vertices r, r+g, … (consecutive 16-byte loads), and the
    // candidates' exact dots are merged: largest first, lowest index among equals (strict '>', src/mesh.cpp:341-358).
    // Correct for any group the hardware happens to form; a group of one is the serial second pass.
    const float thr_own = f1 - 2.0f * E;
    const bool ambiguous = !(f2 < thr_own);
    const unsigned group = __activemask();
    const unsigned todo = __ballot_sync(group, ambiguous);
    if (todo == 0u) return best;
    return hull_exact_pass(vf, v, nverts, lx, ly, lz, thr_own, l.x, l.y, l.z, group, todo, best);
}

// Farthest point of the shape along d.  BIG = false: an instance for contexts that hold no hull above
// HULL_PREFILTER_MIN vertices (the host knows: pk_shape_hull), compiled without the float-prefiltered scan and its
// shared exact pass — sphere / box scenes then do not pay for that code in registers and instruction cache.  (Such
// an instance would still be correct on a big hull: it takes the plain FP64 scan, which is the reference's loop.)
template <bool BIG = true> __device__ __forceinline__ d3 support(const ShapeView &s, d3 d)
{
    if (s.kind == KIND_OBB)
    {
        d3 l = rotate(conjugate(s.q), d);
        d3 sh{(l.x >= 0 ? 1.0 : -1.0) * s.h.x, (l.y >= 0 ? 1.0 : -1.0) * s.h.y, (l.z >= 0 ? 1.0 : -1.0) * s.h.z};
        return s.p + rotate(s.q, sh);
    }
    if (s.kind == KIND_SPHERE) return s.p + s.h.x * normalized(d);
    if (s.kind == KIND_AABB) return {d.x >= 0 ? s.h.x : s.p.x, d.y >= 0 ? s.h.y : s.p.y, d.z >= 0 ? s.h.z : s.p.z};
    // HULL: argmax of v·l over all vertices, strict '>' so the lowest index wins ties
    // (src/mesh.cpp:341-358).  The dot the reference compares is the FP64 value (x·lx + y·ly) + z·lz.
    d3 l = rotate(conjugate(s.q), d);
    if (s.hull_r < 0.f)
    {
        // mesh::box (src/mesh.cpp:31-40): every body of a physkit::world demo is one.  Its eight vertices are
        // (±hx, ±hy, ±hz) in a fixed order, so the eight dots (x·lx + y·ly) + z·lz of the reference's scan share
        // three products and four sums — (−x)·lx = −(x·lx) and (−a) + (−b) = −(a + b) hold exactly under round
        // to nearest — and the scan's strict '>' in index order picks among them.  s.h holds the local minimum.
        const double hx = -s.h.x, hy = -s.h.y, hz = -s.h.z;
        const double px = hx * l.x, py = hy * l.y, pz = hz * l.z;
        const double pp = px + py, pm = px - py;
        const double A = pp + pz, B = pp - pz, C = pm + pz, D = pm - pz;
        // vertex 0 (−,−,−) … 7 (−,+,+)
        const double t[8] = {-A, D, B, -C, -B, C, A, -D};
        int bi = 0;
        double bt = t[0];
#pragma unroll
        for (int i = 1; i < 8; ++i)
        {
            if (t[i] > bt)
            {
                bt = t[i];
                bi = i;
            }
        }
        const double sx = ((0x66 >> bi) & 1) ? hx : -hx; // + for vertices 1, 2, 5, 6
        const double sy = ((0xCC >> bi) & 1) ? hy : -hy; // + for vertices 2, 3, 6, 7
        const double sz = (bi >= 4) ? hz : -hz;
        return rotate(s.q, d3{sx, sy, sz}) + s.p;
    }
    const double *v = s.verts;
    uint32_t best = 0;
    if (!BIG || s.nverts <= HULL_PREFILTER_MIN)
    {
        double best_dot = (v[0] * l.x + v[1] * l.y) + v[2] * l.z;
        for (uint32_t i = 1; i < s.nverts; ++i)
        {
            double t = (v[3 * i] * l.x + v[3 * i + 1] * l.y) + v[3 * i + 2] * l.z;
            if (t > best_dot)
            {
                best_dot = t;
                best = i;
            }
        }
    }
    else
        best = hull_argmax(s.vf, v, s.nverts, s.hull_r, l);
    d3 bv{v[3 * best], v[3 * best + 1], v[3 * best + 2]};
    return rotate(s.q, bv) + s.p;
}

// An element count that lives on the device (a counter an earlier kernel of the step filled), bounded by the number
// the grid was sized for; nullptr: the bound is the count.
__device__ __forceinline__ uint64_t device_count(const unsigned long long *count, uint64_t cap)
{
    if (!count) return cap;
    const unsigned long long v = *count;
    return v < cap ? v : cap;
}

// 88-byte contact record, mirrors pk_contact of include/pk_collide.h.
struct ContactRec
{
    uint64_t key;
    double normal[3];
    double world_a[3];
    double world_b[3];
    double depth;
};
static_assert(sizeof(ContactRec) == 88, "ContactRec must match pk_contact");

} // namespace pk
