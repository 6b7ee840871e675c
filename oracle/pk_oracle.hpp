// pk_oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A plain-double C++20 restatement of the PhysKit collision hot path, written to follow the
// reference operation-for-operation (including Eigen's evaluation order) so that it can act as
// the checker for the CUDA library in physkit_b200/.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may use anything under oracle/.
//
// The reference (C++26 + mp-units 2.5.0 + Eigen 5.0.1 + abseil 20250814.1, conanfile.py:61-79)
// cannot be compiled in this image (g++ 13.3 only, none of the three dependencies on disk), so
// this file is a PORT, pinned against the reference's own known-answer tests
// (tests/gjk/gjk_test.cpp, tests/epa/epa_test.cpp, tests/dynamic_bvh/main.cpp,
// tests/obb/obb_test.cpp, tests/mesh/main.cpp) re-expressed in tests/test_oracle_*.py.
// ulp-level operation order of Eigen is restated from its public headers (un-vendored, so it
// cannot be diffed here): "parity unpinned" below 1e-9 absolute — see DESIGN.md §oracle.
//
// Compile with:  g++ -O3 -std=c++20 -ffp-contract=off   (no -march, no fast-math; mirrors the
// reference's Release flags, CMakeLists.txt:126-132 → SSE2, no FMA).
//
// All citations are relative to /root/reference/.
#pragma once

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <optional>
#include <unordered_set>
#include <utility>
#include <vector>

namespace pko
{

// ---------------------------------------------------------------------------------------------
// Algebra: Eigen fixed-size double semantics (include/physkit/algebra/lin_alg.h wraps Eigen).
// ---------------------------------------------------------------------------------------------
struct v3
{
    double x, y, z;
};

inline v3 operator+(v3 a, v3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline v3 operator-(v3 a, v3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline v3 operator-(v3 a) { return {-a.x, -a.y, -a.z}; }
inline v3 operator*(v3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline v3 operator*(double s, v3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline v3 operator/(v3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }

// Eigen redux over a fixed 3-vector of double with SSE2 packets (PacketSize 2, unaligned
// vectorisation allowed): predux(packet{x0,x1}) then the scalar tail  →  (x0 + x1) + x2.
// lin_alg.h:212-213 (dot), :229 (squared_norm).
inline double dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline double sqnorm(v3 a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }
inline double norm(v3 a) { return std::sqrt(sqnorm(a)); }

// Eigen cross3: (a1*b2 - a2*b1, a2*b0 - a0*b2, a0*b1 - a1*b0).  lin_alg.h:215-219.
inline v3 cross(v3 a, v3 b)
{
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Eigen MatrixBase::normalized(): z = squaredNorm(); z > 0 ? v / sqrt(z) : v  (true division).
// lin_alg.h:232-240.
inline v3 normalized(v3 a)
{
    double z = sqnorm(a);
    if (z > 0.0)
    {
        double s = std::sqrt(z);
        return {a.x / s, a.y / s, a.z / s};
    }
    return a;
}

inline v3 cwise(v3 a, v3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }

// Eigen coefficient order x,y,z,w (lin_alg.h:388); ctor order is (w,x,y,z) (lin_alg.h:428-432).
struct quat
{
    double x, y, z, w;
};

inline quat conjugate(quat q) { return {-q.x, -q.y, -q.z, q.w}; }

// Eigen QuaternionBase::_transformVector (lin_alg.h:493-499):
//   uv = vec().cross(v); uv += uv; return v + w()*uv + vec().cross(uv);   (left-assoc sum)
inline v3 rotate(quat q, v3 v)
{
    v3 qv{q.x, q.y, q.z};
    v3 uv = cross(qv, v);
    uv = uv + uv;
    v3 c = cross(qv, uv);
    return {(v.x + q.w * uv.x) + c.x, (v.y + q.w * uv.y) + c.y, (v.z + q.w * uv.z) + c.z};
}

// ---------------------------------------------------------------------------------------------
// aabb — include/physkit/collision/bounds.h:27-205
// ---------------------------------------------------------------------------------------------
struct aabb
{
    v3 min, max;

    // bounds.h:65-71
    v3 point(unsigned i) const
    {
        return {(i & 1u) ? max.x : min.x, (i & 2u) ? max.y : min.y, (i & 4u) ? max.z : min.z};
    }
    // bounds.h:80-85
    bool contains(const aabb &o) const
    {
        return (o.min.x >= min.x && o.max.x <= max.x) && (o.min.y >= min.y && o.max.y <= max.y) &&
               (o.min.z >= min.z && o.max.z <= max.z);
    }
    // bounds.h:87-92 (inclusive)
    bool intersects(const aabb &o) const
    {
        return (min.x <= o.max.x && max.x >= o.min.x) && (min.y <= o.max.y && max.y >= o.min.y) &&
               (min.z <= o.max.z && max.z >= o.min.z);
    }
    // bounds.h:59-63
    double surface_area() const
    {
        v3 s = max - min;
        return 2.0 * ((s.x * s.y + s.y * s.z) + s.z * s.x);
    }
    // bounds.h:164-174
    v3 support(v3 d) const
    {
        return {d.x >= 0 ? max.x : min.x, d.y >= 0 ? max.y : min.y, d.z >= 0 ? max.z : min.z};
    }
};

// ---------------------------------------------------------------------------------------------
// ray — include/physkit/collision/bvh.h:22-98 (slab test only; the triangle test is not on the path)
// ---------------------------------------------------------------------------------------------
inline double safe_inv(double x) { return 1.0 / x; } // bvh.h:27-35, IEC 559 branch: ±0 → ±inf

struct ray
{
    v3 origin, direction;
    ray(v3 o, v3 d) : origin(o), direction(normalized(d)) {} // bvh.h:47-50

    // bvh.h:59-98.  0·inf = NaN is replaced by the limit (−inf for the near plane, +inf for the far one).
    std::optional<double> intersect_distance(const aabb &box, double max_distance) const
    {
        auto slab = [](double lo, double hi)
        {
            if (std::isnan(lo)) lo = -std::numeric_limits<double>::infinity();
            if (std::isnan(hi)) hi = std::numeric_limits<double>::infinity();
            return std::pair{std::fmin(lo, hi), std::fmax(lo, hi)};
        };
        auto [tminx, tmaxx] = slab((box.min.x - origin.x) * safe_inv(direction.x), (box.max.x - origin.x) * safe_inv(direction.x));
        auto [tminy, tmaxy] = slab((box.min.y - origin.y) * safe_inv(direction.y), (box.max.y - origin.y) * safe_inv(direction.y));
        auto [tminz, tmaxz] = slab((box.min.z - origin.z) * safe_inv(direction.z), (box.max.z - origin.z) * safe_inv(direction.z));
        double tmin = std::max({tminx, tminy, tminz});
        double tmax = std::min({tmaxx, tmaxy, tmaxz});
        if (tmax >= 0.0 && tmin <= tmax && tmin <= max_distance) return std::max(0.0, tmin);
        return std::nullopt;
    }
};

// bounds.h:94-102
inline aabb aabb_union(const aabb &a, const aabb &b)
{
    return {{std::min(a.min.x, b.min.x), std::min(a.min.y, b.min.y), std::min(a.min.z, b.min.z)},
            {std::max(a.max.x, b.max.x), std::max(a.max.y, b.max.y), std::max(a.max.z, b.max.z)}};
}

// bounds.h:33-49
inline aabb aabb_from_points(const v3 *pts, std::size_t n)
{
    aabb box{pts[0], pts[0]};
    for (std::size_t i = 1; i < n; ++i)
    {
        box.min.x = std::min(box.min.x, pts[i].x);
        box.min.y = std::min(box.min.y, pts[i].y);
        box.min.z = std::min(box.min.z, pts[i].z);
        box.max.x = std::max(box.max.x, pts[i].x);
        box.max.y = std::max(box.max.y, pts[i].y);
        box.max.z = std::max(box.max.z, pts[i].z);
    }
    return box;
}

// aabb::operator*(unit_quat) bounds.h:142-158 followed by operator+(offset) bounds.h:112-113;
// together this is mesh::instance::bounds()  src/mesh.cpp:398.
inline aabb instance_bounds(const aabb &local, v3 pos, quat q)
{
    v3 p0 = rotate(q, local.min);
    aabb r{p0, p0};
    for (unsigned i = 1; i < 8; ++i)
    {
        v3 pt = rotate(q, local.point(i));
        r.min.x = std::min(r.min.x, pt.x);
        r.min.y = std::min(r.min.y, pt.y);
        r.min.z = std::min(r.min.z, pt.z);
        r.max.x = std::max(r.max.x, pt.x);
        r.max.y = std::max(r.max.y, pt.y);
        r.max.z = std::max(r.max.z, pt.z);
    }
    return {r.min + pos, r.max + pos};
}

// dynamic_bvh::update_leaf fat rule  src/bvh.cpp:483-508.  Returns true when the leaf is re-inserted.
inline bool fat_update(aabb &stored, const aabb &true_bounds, v3 disp)
{
    if (stored.contains(true_bounds)) return false;
    const double margin = .1;
    v3 mv{margin, margin, margin};
    aabb nb{true_bounds.min - mv, true_bounds.max + mv};
    if (disp.x < 0.0) nb.min.x = nb.min.x + disp.x; else nb.max.x = nb.max.x + disp.x;
    if (disp.y < 0.0) nb.min.y = nb.min.y + disp.y; else nb.max.y = nb.max.y + disp.y;
    if (disp.z < 0.0) nb.min.z = nb.min.z + disp.z; else nb.max.z = nb.max.z + disp.z;
    stored = nb;
    return true;
}

// ---------------------------------------------------------------------------------------------
// Support shapes — bounds.h:164-174 (aabb), :539-548 (obb), :328-329 (bounding_sphere),
// src/mesh.cpp:341-358 (mesh::support), :442-448 (mesh::instance::support).
// ---------------------------------------------------------------------------------------------
enum shape_kind : int32_t
{
    KIND_AABB = 0,
    KIND_OBB = 1,
    KIND_SPHERE = 2,
    KIND_HULL = 3,
};

struct shape
{
    int32_t kind;
    v3 a;            // AABB: min | OBB: center | SPHERE: center | HULL: position
    v3 b;            // AABB: max | OBB: half extents | SPHERE: (r,-,-)
    quat q;          // OBB / HULL orientation
    const v3 *verts; // HULL
    uint32_t nverts;
};

inline v3 support(const shape &s, v3 d)
{
    switch (s.kind)
    {
    case KIND_AABB:
        return aabb{s.a, s.b}.support(d);
    case KIND_OBB:
    {
        v3 l = rotate(conjugate(s.q), d);
        v3 signs{l.x >= 0 ? 1.0 : -1.0, l.y >= 0 ? 1.0 : -1.0, l.z >= 0 ? 1.0 : -1.0};
        return s.a + rotate(s.q, cwise(signs, s.b));
    }
    case KIND_SPHERE:
        return s.a + s.b.x * normalized(d);
    default:
    {
        v3 l = rotate(conjugate(s.q), d);
        std::size_t best = 0;
        double best_dot = dot(s.verts[0], l);
        for (std::size_t i = 1; i < s.nverts; ++i)
        {
            double v = dot(s.verts[i], l);
            if (v > best_dot)
            {
                best_dot = v;
                best = i;
            }
        }
        return rotate(s.q, s.verts[best]) + s.a;
    }
    }
}

// collision.h:27-49
struct support_pt
{
    v3 p, pa, pb;
};

inline support_pt minkowski_support(const shape &a, const shape &b, v3 d)
{
    v3 pa = support(a, d);
    v3 pb = support(b, -d);
    return {pa - pb, pa, pb};
}

// collision.h:52-59
struct collision_info
{
    v3 normal, world_a, world_b;
    double depth;
};

// ---------------------------------------------------------------------------------------------
// GJK — src/collision.cpp:10-189
// ---------------------------------------------------------------------------------------------
struct simplex_t
{
    support_pt pts[4];
    int n = 0;
    void push_back(const support_pt &p) { pts[n++] = p; }
    void erase(int i)
    {
        for (int k = i; k + 1 < n; ++k) pts[k] = pts[k + 1];
        --n;
    }
    support_pt &operator[](int i) { return pts[i]; }
    const support_pt &operator[](int i) const { return pts[i]; }
    int size() const { return n; }
};

// collision.cpp:12-41
inline bool handle_line(simplex_t &s, v3 &direction)
{
    constexpr double eps = 1e-12;
    const support_pt a = s[1];
    const support_pt b = s[0];
    const v3 ab = b.p - a.p;
    const v3 ao = -a.p;
    double ab_dot_ao = dot(ab, ao);
    if (ab_dot_ao > 0.0)
    {
        v3 triple = cross(cross(ab, ao), ab);
        if (sqnorm(triple) < eps)
        {
            v3 ab_hat = normalized(ab);
            v3 perp = cross(ab_hat, v3{0.0, 1.0, 0.0});
            if (sqnorm(perp) < eps) perp = cross(ab_hat, v3{0.0, 0.0, 1.0});
            direction = normalized(perp);
        }
        else
            direction = normalized(triple);
    }
    else
    {
        s.erase(0);
        direction = normalized(ao);
    }
    return false;
}

// collision.cpp:43-88
inline bool handle_triangle(simplex_t &s, v3 &direction)
{
    const support_pt a = s[2];
    const support_pt b = s[1];
    const support_pt c = s[0];
    const v3 ab = b.p - a.p;
    const v3 ac = c.p - a.p;
    const v3 ao = -a.p;
    const v3 abc = cross(ab, ac);

    const v3 ab_perp = cross(ab, abc);
    if (dot(ab_perp, ao) > 0.0)
    {
        s.erase(0);
        v3 triple = cross(cross(ab, ao), ab);
        direction = (sqnorm(triple) < 1e-12) ? normalized(ao) : normalized(triple);
        return false;
    }
    const v3 ac_perp = cross(abc, ac);
    if (dot(ac_perp, ao) > 0.0)
    {
        s.erase(1);
        v3 triple = cross(cross(ac, ao), ac);
        direction = (sqnorm(triple) < 1e-12) ? normalized(ao) : normalized(triple);
        return false;
    }
    if (dot(abc, ao) <= 0.0)
    {
        std::swap(s[0], s[1]);
        direction = normalized(-abc);
    }
    else
        direction = normalized(abc);
    return false;
}

// collision.cpp:90-147
inline bool handle_tetrahedron(simplex_t &s, v3 &direction)
{
    const support_pt a = s[3];
    const support_pt b = s[2];
    const support_pt c = s[1];
    const support_pt d = s[0];
    const v3 ao = -a.p;

    v3 abc = cross(b.p - a.p, c.p - a.p);
    v3 acd = cross(c.p - a.p, d.p - a.p);
    v3 adb = cross(d.p - a.p, b.p - a.p);

    auto orient = [](v3 &n, const support_pt &fa, const support_pt &opp)
    {
        v3 t = opp.p - fa.p;
        if (dot(n, t) > 0.0) n = -n;
    };
    orient(abc, a, d);
    orient(acd, a, b);
    orient(adb, a, c);

    if (dot(abc, ao) > 0.0)
    {
        s.n = 0;
        s.push_back(c);
        s.push_back(b);
        s.push_back(a);
        direction = normalized(abc);
        return handle_triangle(s, direction);
    }
    if (dot(acd, ao) > 0.0)
    {
        s.n = 0;
        s.push_back(d);
        s.push_back(c);
        s.push_back(a);
        direction = normalized(acd);
        return handle_triangle(s, direction);
    }
    if (dot(adb, ao) > 0.0)
    {
        s.n = 0;
        s.push_back(b);
        s.push_back(d);
        s.push_back(a);
        direction = normalized(adb);
        return handle_triangle(s, direction);
    }
    return true;
}

// collision.cpp:149-162
inline bool handle_simplex(simplex_t &s, v3 &direction)
{
    switch (s.size())
    {
    case 2: return handle_line(s, direction);
    case 3: return handle_triangle(s, direction);
    case 4: return handle_tetrahedron(s, direction);
    default: return false;
    }
}

struct gjk_stats
{
    int gjk_iters = 0;
    int epa_iters = 0;
    int epa_faces = 0;
    int epa_verts = 0;
    int epa_heap_max = 0;
    int epa_horizon_max = 0;
    int epa_stack_max = 0;
    int exit_code = 0; // 0 miss, 1 converged, 2 best-guess, 3 pad fail, 4 heap empty
};

// collision.cpp:165-189
inline std::optional<simplex_t> gjk_collision(const shape &a, const shape &b, gjk_stats *st = nullptr)
{
    constexpr double eps = 1e-12;
    simplex_t s;
    v3 direction{1.0, 0.0, 0.0};

    support_pt point = minkowski_support(a, b, direction);
    s.push_back(point);
    if (sqnorm(point.p) < eps) return s;
    direction = -normalized(point.p);

    constexpr int max_iterations = 100;
    for (int iter = 0; iter < max_iterations; ++iter)
    {
        if (st) st->gjk_iters = iter + 1;
        support_pt np = minkowski_support(a, b, direction);
        double progress = dot(np.p, direction);
        if (progress <= 0) return std::nullopt;
        s.push_back(np);
        if (handle_simplex(s, direction)) return s;
    }
    return std::nullopt;
}

// collision.cpp:191-248
inline bool pad_simplex(const shape &a, const shape &b, simplex_t &s)
{
    switch (s.size())
    {
    case 1:
    {
        v3 dir{1, 0, 0};
        support_pt p2 = minkowski_support(a, b, dir);
        if (sqnorm(p2.p - s[0].p) < 1e-6) p2 = minkowski_support(a, b, -dir);
        s.push_back(p2);
    }
        [[fallthrough]];
    case 2:
    {
        v3 line = s[1].p - s[0].p;
        v3 dir = cross(normalized(line), v3{0.0, 1.0, 0.0});
        if (sqnorm(dir) < 1e-6) dir = cross(normalized(line), v3{0.0, 0.0, 1.0});
        dir = normalized(dir); // Eigen normalize(): in-place /= sqrt(z) when z > 0
        support_pt p3 = minkowski_support(a, b, dir);
        if (sqnorm(cross(line, p3.p - s[0].p)) < 1e-6) p3 = minkowski_support(a, b, -dir);
        s.push_back(p3);
    }
        [[fallthrough]];
    case 3:
    {
        v3 ab = s[1].p - s[0].p;
        v3 ac = s[2].p - s[0].p;
        v3 dir = normalized(cross(ab, ac));
        support_pt p4 = minkowski_support(a, b, dir);
        if (std::abs(dot(p4.p - s[0].p, dir)) < 1e-6) p4 = minkowski_support(a, b, -dir);
        s.push_back(p4);
    }
    default:
        break;
    }
    v3 ad = s[0].p - s[3].p;
    v3 bd = s[1].p - s[3].p;
    v3 cd = s[2].p - s[3].p;
    double triple = dot(ad, cross(bd, cd));
    return std::abs(triple) > 1e-12;
}

// ---------------------------------------------------------------------------------------------
// EPA — src/collision.cpp:251-509.  The reference keeps faces / polytope / heap in
// absl::InlinedVector<…,128> that spill to the heap beyond 128 entries; std::vector reproduces
// that (unbounded).  The face heap uses std::ranges::push_heap / pop_heap with the comparator
// faces[f1].distance > faces[f2].distance; here the libstdc++ std::push_heap / std::pop_heap
// (identical algorithm) are used.  Tie order among equal distances is therefore libstdc++'s.
// ---------------------------------------------------------------------------------------------
struct epa_solver
{
    using index_t = std::uint16_t;
    static constexpr index_t null_index = static_cast<index_t>(-1);

    struct face
    {
        std::array<index_t, 3> vertices{};
        std::array<index_t, 3> adj = {null_index, null_index, null_index};
        v3 normal{};
        double distance{};
        bool obsolete = false;
    };
    struct silhouette_edge
    {
        index_t start_vertex, end_vertex, adjacent_face;
    };

    std::vector<face> faces;
    std::vector<support_pt> polytope;
    std::vector<std::size_t> face_heap;
    int stack_max = 0;
    int heap_max = 0;

    // collision.cpp:273-297
    void init_face(std::size_t f_idx, std::size_t i, std::size_t j, std::size_t k,
                   index_t opposite_index = null_index)
    {
        face &f = faces[f_idx];
        f.vertices = {static_cast<index_t>(i), static_cast<index_t>(j), static_cast<index_t>(k)};
        v3 ab = polytope[j].p - polytope[i].p;
        v3 ac = polytope[k].p - polytope[i].p;
        f.normal = cross(ab, ac);
        if (sqnorm(f.normal) < 1e-12)
            f.normal = v3{0, 0, 0};
        else
            f.normal = normalized(f.normal); // Eigen normalize(): z>0 ? /= sqrt(z)
        if (opposite_index != null_index &&
            dot(f.normal, polytope[opposite_index].p - polytope[i].p) > 0.0)
        {
            std::swap(f.vertices[1], f.vertices[2]);
            f.normal = -f.normal;
        }
        f.distance = dot(f.normal, polytope[i].p);
    }

    std::size_t allocate_face()
    {
        faces.emplace_back();
        return faces.size() - 1;
    }

    // collision.cpp:305-313
    void link_faces(std::size_t f1, std::size_t f2, std::size_t v_a, std::size_t v_b)
    {
        face &a = faces[f1];
        face &b = faces[f2];
        int e1 = (a.vertices[0] == v_a) ? 0 : (a.vertices[1] == v_a ? 1 : 2);
        int e2 = (b.vertices[0] == v_b) ? 0 : (b.vertices[1] == v_b ? 1 : 2);
        a.adj[e1] = static_cast<index_t>(f2);
        b.adj[e2] = static_cast<index_t>(f1);
    }

    // collision.cpp:315-353
    std::vector<silhouette_edge> find_silhouette(index_t face_idx, v3 p)
    {
        constexpr double tolerance = 1e-6;
        std::vector<std::size_t> stack{face_idx};
        faces[face_idx].obsolete = true;
        std::vector<silhouette_edge> horizon;
        while (!stack.empty())
        {
            stack_max = std::max<int>(stack_max, static_cast<int>(stack.size()));
            std::size_t cur = stack.back();
            stack.pop_back();
            face &cur_face = faces[cur];
            for (std::size_t i = 0; i < 3; ++i)
            {
                std::size_t n_idx = cur_face.adj[i];
                if (n_idx == null_index) continue;
                face &nb = faces[n_idx];
                if (nb.obsolete) continue;
                if (dot(nb.normal, p) > nb.distance + tolerance)
                {
                    nb.obsolete = true;
                    stack.push_back(n_idx);
                }
                else
                    horizon.push_back({cur_face.vertices[i], cur_face.vertices[(i + 1) % 3],
                                       static_cast<index_t>(n_idx)});
            }
        }
        return horizon;
    }

    // collision.cpp:390-408
    void push_face(std::size_t f_idx)
    {
        face_heap.push_back(f_idx);
        std::push_heap(face_heap.begin(), face_heap.end(), [&](std::size_t f1, std::size_t f2)
                       { return faces[f1].distance > faces[f2].distance; });
        heap_max = std::max<int>(heap_max, static_cast<int>(face_heap.size()));
    }
    std::size_t pop_face()
    {
        while (!face_heap.empty())
        {
            std::pop_heap(face_heap.begin(), face_heap.end(), [&](std::size_t f1, std::size_t f2)
                          { return faces[f1].distance > faces[f2].distance; });
            std::size_t f_idx = face_heap.back();
            face_heap.pop_back();
            if (!faces[f_idx].obsolete) return f_idx;
        }
        return static_cast<std::size_t>(null_index);
    }

    // collision.cpp:355-388
    void build_initial_tetrahedron()
    {
        auto aip = [&](std::size_t i, std::size_t j, std::size_t k, index_t opp)
        {
            std::size_t f = allocate_face();
            init_face(f, i, j, k, opp);
            push_face(f);
        };
        aip(0, 1, 2, 3);
        aip(0, 2, 3, 1);
        aip(0, 3, 1, 2);
        aip(1, 3, 2, 0);
        for (std::size_t i = 0; i < 4; ++i)
            for (std::size_t j = i + 1; j < 4; ++j)
                for (std::size_t e1 = 0; e1 < 3; ++e1)
                {
                    auto u1 = faces[i].vertices[e1];
                    auto v1 = faces[i].vertices[(e1 + 1) % 3];
                    for (std::size_t e2 = 0; e2 < 3; ++e2)
                    {
                        auto u2 = faces[j].vertices[e2];
                        auto v2 = faces[j].vertices[(e2 + 1) % 3];
                        if (u1 == v2 && v1 == u2)
                        {
                            faces[i].adj[e1] = static_cast<index_t>(j);
                            faces[j].adj[e2] = static_cast<index_t>(i);
                        }
                    }
                }
    }

    // collision.cpp:424-454
    collision_info get_barycentric(const face &f) const
    {
        const support_pt &p0 = polytope[f.vertices[0]];
        const support_pt &p1 = polytope[f.vertices[1]];
        const support_pt &p2 = polytope[f.vertices[2]];
        v3 pm = f.normal * f.distance;
        v3 v0 = p1.p - p0.p;
        v3 v1 = p2.p - p0.p;
        v3 v2 = pm - p0.p;
        double d00 = dot(v0, v0);
        double d01 = dot(v0, v1);
        double d11 = dot(v1, v1);
        double d20 = dot(v2, v0);
        double d21 = dot(v2, v1);
        double denom = d00 * d11 - d01 * d01;
        double v = (d11 * d20 - d01 * d21) / denom;
        double w = (d00 * d21 - d01 * d20) / denom;
        double u = 1.0 - v - w;
        return {-f.normal, (u * p0.pa + v * p1.pa) + w * p2.pa, (u * p0.pb + v * p1.pb) + w * p2.pb,
                f.distance};
    }

    // collision.cpp:411-504
    static std::optional<collision_info> solve(const shape &a, const shape &b, simplex_t &s,
                                               gjk_stats *st = nullptr)
    {
        if (s.size() < 4 && !pad_simplex(a, b, s))
        {
            if (st) st->exit_code = 3;
            return std::nullopt;
        }
        epa_solver solver;
        solver.polytope.assign(s.pts, s.pts + s.n);
        solver.build_initial_tetrahedron();

        constexpr int max_iterations = 64;
        constexpr double tolerance = 1e-6;
        int horizon_max = 0;
        auto fill_stats = [&](int code, int iters)
        {
            if (!st) return;
            st->exit_code = code;
            st->epa_iters = iters;
            st->epa_faces = static_cast<int>(solver.faces.size());
            st->epa_verts = static_cast<int>(solver.polytope.size());
            st->epa_heap_max = solver.heap_max;
            st->epa_horizon_max = horizon_max;
            st->epa_stack_max = solver.stack_max;
        };

        int iter = 0;
        for (; iter < max_iterations; ++iter)
        {
            std::size_t min_idx = solver.pop_face();
            if (min_idx == null_index) break;
            // (the reference holds a reference to faces[min_idx]; it is only read before any
            //  allocate_face(), so a copy of normal/distance is equivalent)
            const v3 mn = solver.faces[min_idx].normal;
            const double md = solver.faces[min_idx].distance;

            support_pt p = minkowski_support(a, b, mn);
            double p_dist = dot(mn, p.p);
            if (p_dist - md < tolerance)
            {
                fill_stats(1, iter + 1);
                return solver.get_barycentric(solver.faces[min_idx]);
            }
            auto horizon = solver.find_silhouette(static_cast<index_t>(min_idx), p.p);
            horizon_max = std::max<int>(horizon_max, static_cast<int>(horizon.size()));
            if (horizon.empty()) break;

            solver.polytope.push_back(p);
            index_t p_idx = static_cast<index_t>(solver.polytope.size() - 1);

            std::vector<std::size_t> new_faces;
            for (const auto &[start, end, adj_face] : horizon)
            {
                std::size_t f = solver.allocate_face();
                solver.init_face(f, start, end, p_idx);
                solver.link_faces(f, adj_face, start, end);
                solver.push_face(f);
                new_faces.push_back(f);
            }
            std::size_t n = new_faces.size();
            for (std::size_t i = 0; i < n; ++i)
                for (std::size_t j = i + 1; j < n; ++j)
                {
                    if (horizon[i].end_vertex == horizon[j].start_vertex)
                        solver.link_faces(new_faces[i], new_faces[j], horizon[i].end_vertex, p_idx);
                    else if (horizon[i].start_vertex == horizon[j].end_vertex)
                        solver.link_faces(new_faces[j], new_faces[i], horizon[j].end_vertex, p_idx);
                }
        }
        std::size_t min_idx = solver.pop_face();
        if (min_idx == null_index)
        {
            fill_stats(4, iter);
            return std::nullopt;
        }
        fill_stats(2, iter);
        return solver.get_barycentric(solver.faces[min_idx]);
    }
};

// ---------------------------------------------------------------------------------------------
// manifold / narrow_phase::calculate — include/physkit/collision/collision_phases.h:75-320
// (the consumer of gjk_epa's result; restated for the device-side manifold update)
// ---------------------------------------------------------------------------------------------
struct contact_info_t // manifold::contact_info (:93-99) with contact_point (:75-88) flattened
{
    v3 normal{0, 0, 0}, local_a{0, 0, 0}, local_b{0, 0, 0};
    double depth = 0.0;
    double normal_impulse = 0.0;
    double tangent_impulses[2] = {0.0, 0.0};
};

struct manifold_t
{
    static constexpr int max_contact_points = 4; // :101
    contact_info_t c[4];
    int n = 0;

    // :127-133
    void add_contact(const contact_info_t &info)
    {
        if (n < max_contact_points)
            c[n++] = info;
        else
            add_reduce(info);
    }
    // :139-198 — keep the deepest point, the one farthest from it, the one spanning the largest triangle
    // with those two, and the one farthest from the third
    void add_reduce(const contact_info_t &new_pt)
    {
        contact_info_t pool[5];
        for (int i = 0; i < 4; ++i) pool[i] = c[i];
        pool[4] = new_pt;
        int best[4] = {0, 1, 2, 3};
        for (int i = 1; i < 5; ++i)
            if (pool[i].depth > pool[best[0]].depth) best[0] = i;
        double max_dist2 = -1.0;
        for (int i = 0; i < 5; ++i)
        {
            if (i == best[0]) continue;
            double dist2 = sqnorm(pool[i].local_a - pool[best[0]].local_a);
            if (dist2 > max_dist2)
            {
                max_dist2 = dist2;
                best[1] = i;
            }
        }
        double max_area2 = -1.0;
        v3 edge0 = pool[best[1]].local_a - pool[best[0]].local_a;
        for (int i = 0; i < 5; ++i)
        {
            if (i == best[0] || i == best[1]) continue;
            v3 edge1 = pool[i].local_a - pool[best[0]].local_a;
            double area2 = sqnorm(cross(edge0, edge1));
            if (area2 > max_area2)
            {
                max_area2 = area2;
                best[2] = i;
            }
        }
        max_dist2 = -1.0;
        for (int i = 0; i < 5; ++i)
        {
            if (i == best[0] || i == best[1] || i == best[2]) continue;
            double dist2 = sqnorm(pool[i].local_a - pool[best[2]].local_a);
            if (dist2 > max_dist2)
            {
                max_dist2 = dist2;
                best[3] = i;
            }
        }
        n = 0;
        for (int b : best) c[n++] = pool[b];
    }
};

// One pair's share of narrow_phase::calculate (:252-318): merge this step's contact (if any) with the
// manifold kept from the previous step.  pos/q of the two bodies as of this step (particle.h:107-111).
inline manifold_t manifold_merge(const manifold_t &old_man, const contact_info_t *new_contact_in, v3 pos_a, quat q_a, v3 pos_b,
                                 quat q_b)
{
    constexpr double distance2_eps = .005 * .005;         // :212
    constexpr double contact_breaking_threshold = 0.05;  // :213
    constexpr double drift2_eps = 0.06;                  // :267
    contact_info_t nc;
    const bool have_new = new_contact_in != nullptr;
    if (have_new) nc = *new_contact_in;
    manifold_t new_man;
    for (int k = 0; k < old_man.n; ++k)
    {
        contact_info_t old_contact = old_man.c[k];
        if (have_new)
        {
            // warm start the new point from an old one at (nearly) the same place on either body
            if (sqnorm(nc.local_a - old_contact.local_a) < distance2_eps || sqnorm(nc.local_b - old_contact.local_b) < distance2_eps)
            {
                nc.normal_impulse = old_contact.normal_impulse;
                nc.tangent_impulses[0] = old_contact.tangent_impulses[0];
                nc.tangent_impulses[1] = old_contact.tangent_impulses[1];
                continue;
            }
        }
        v3 world_old_a = rotate(q_a, old_contact.local_a) + pos_a; // project_to_world
        v3 world_old_b = rotate(q_b, old_contact.local_b) + pos_b;
        v3 normal = have_new ? nc.normal : old_contact.normal;
        v3 relative = world_old_b - world_old_a;
        double depth = dot(relative, normal);
        v3 projected_a = world_old_a + normal * depth;
        double drift2 = sqnorm(projected_a - world_old_b);
        if (depth > -contact_breaking_threshold && drift2 < drift2_eps)
        {
            old_contact.depth = depth;
            old_contact.normal = normal;
            new_man.add_contact(old_contact);
        }
    }
    if (have_new) new_man.add_contact(nc);
    return new_man;
}

// ---------------------------------------------------------------------------------------------
// Rigid-body state and the integrator — include/physkit/core/particle.h:12-150,
// include/physkit/detail/integrate.h:17-47, the per-body loops of src/world.cpp:22-34, 50-55.
// Eigen's fixed-size 3×3 kernels are restated as coefficient sums in k order, (k0 + k1) + k2, like dot();
// the quaternion product and sin / cos are "parity unpinned" at the ulp level (Eigen has a SIMD path for
// double quaternions; libm vs libdevice): tests compare these with a stated tolerance.
// ---------------------------------------------------------------------------------------------
struct m3
{
    double m[3][3]; // [row][col]
};
inline m3 mul(const m3 &a, const m3 &b)
{
    m3 c;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c.m[i][j] = (a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j]) + a.m[i][2] * b.m[2][j];
    return c;
}
inline v3 mul(const m3 &a, v3 v)
{
    return {(a.m[0][0] * v.x + a.m[0][1] * v.y) + a.m[0][2] * v.z, (a.m[1][0] * v.x + a.m[1][1] * v.y) + a.m[1][2] * v.z,
            (a.m[2][0] * v.x + a.m[2][1] * v.y) + a.m[2][2] * v.z};
}
inline m3 transpose(const m3 &a)
{
    m3 t;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) t.m[i][j] = a.m[j][i];
    return t;
}
// Eigen compute_inverse_size3_helper: cofactors of column 0, determinant = their dot with column 0,
// every entry = cofactor · (1 / det)   (lin_alg.h mat3::inverse → Eigen::Matrix3d::inverse)
inline m3 inverse(const m3 &a)
{
    auto cof = [&](int i, int j)
    {
        int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return a.m[i1][j1] * a.m[i2][j2] - a.m[i1][j2] * a.m[i2][j1];
    };
    double c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    double det = (c00 * a.m[0][0] + c10 * a.m[1][0]) + c20 * a.m[2][0];
    double invdet = 1.0 / det;
    m3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[j][i] = cof(i, j) * invdet;
    return r;
}
// Eigen QuaternionBase::toRotationMatrix (lin_alg.h:530-535)
inline m3 to_rotation_matrix(quat q)
{
    double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
    double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    m3 r;
    r.m[0][0] = 1.0 - (tyy + tzz);
    r.m[0][1] = txy - twz;
    r.m[0][2] = txz + twy;
    r.m[1][0] = txy + twz;
    r.m[1][1] = 1.0 - (txx + tzz);
    r.m[1][2] = tyz - twx;
    r.m[2][0] = txz - twy;
    r.m[2][1] = tyz + twx;
    r.m[2][2] = 1.0 - (txx + tyy);
    return r;
}
// Eigen quat_product (scalar form; lin_alg.h:478-483)
inline quat qmul(quat a, quat b)
{
    return {((a.w * b.x + a.x * b.w) + a.y * b.z) - a.z * b.y, ((a.w * b.y + a.y * b.w) + a.z * b.x) - a.x * b.z,
            ((a.w * b.z + a.z * b.w) + a.x * b.y) - a.y * b.x, ((a.w * b.w - a.x * b.x) - a.y * b.y) - a.z * b.z};
}
inline quat qnormalized(quat q)
{
    double n = std::sqrt(((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w);
    return {q.x / n, q.y / n, q.z / n, q.w / n};
}
// detail::exp (integrate.h:21-32)
inline quat exp_rotation(v3 ang_vel, double dt)
{
    v3 angle = ang_vel * dt;
    double mag = norm(angle);
    if (mag < 1e-12)
    {
        v3 half = angle * 0.5;
        return qnormalized(quat{half.x, half.y, half.z, 1.0});
    }
    v3 axis = angle / mag;
    double s = std::sin(0.5 * mag), c = std::cos(0.5 * mag); // Eigen AngleAxis → Quaternion (lin_alg.h:569-576)
    return {s * axis.x, s * axis.y, s * axis.z, c};
}

struct rigid_state // particle.h:113-139
{
    v3 pos{}, vel{}, acc{}, ang_vel{}, torque{};
    quat q{0, 0, 0, 1};
    double mass = 1.0, inv_mass = 1.0;
    m3 inertia_local{}, inv_inertia_local{};
    m3 inertia_world{}, inv_inertia_world{};
    bool is_static = false;

    void update_derived_state() // particle.h:140-146
    {
        m3 r = to_rotation_matrix(q);
        m3 rt = transpose(r);
        inertia_world = mul(mul(r, inertia_local), rt);
        inv_inertia_world = mul(mul(r, inv_inertia_local), rt);
    }
    v3 angular_accel() const // particle.h:72-76
    {
        return mul(inv_inertia_world, torque - cross(ang_vel, mul(inertia_world, ang_vel)));
    }
};

// world::step_impl loop A without the broad-phase call (src/world.cpp:22-34): returns vel·dt, the
// displacement handed to update_node
inline v3 step_velocity(rigid_state &o, v3 gravity, double dt)
{
    o.acc = o.acc + (gravity * o.mass) * o.inv_mass;        // apply_force(gravity * mass), particle.h:78-79
    o.vel = o.vel + o.acc * dt;                             // integrate_vel, integrate.h:36-40
    o.ang_vel = o.ang_vel + o.angular_accel() * dt;
    v3 disp = o.vel * dt;
    o.acc = v3{0, 0, 0};                                    // clear_forces, particle.h:101-105
    o.torque = v3{0, 0, 0};
    return disp;
}
// loop B (src/world.cpp:50-55): integrate_pos, integrate.h:42-46
inline void step_position(rigid_state &o, double dt)
{
    o.pos = o.pos + o.vel * dt;
    o.q = qmul(exp_rotation(o.ang_vel, dt), o.q);
    o.update_derived_state();
}

// ---------------------------------------------------------------------------------------------
// Contact Jacobians — include/physkit/collision/constraint.h:104-113 (build_orthonormal_basis), :874-953
// (build_contact_jacobian), :1052-1104 (setup_contacts).  SURVEY §8 f2.
// ---------------------------------------------------------------------------------------------
struct jacobian_row_t // constraint.h:29-47, the fields setup_contacts fills
{
    v3 J_v{}, J_w_a{}, J_w_b{};
    double M_eff = 0.0, bias = 0.0;
};
struct contact_solver_point_t // constraint.h:1204-1214 + the warm-start impulses (:1096-1098)
{
    jacobian_row_t normal, tangent1, tangent2;
    double friction_coeff = 0.0, inv_m_11 = 0.0, inv_m_12 = 0.0, inv_m_22 = 0.0;
    double accumulated[3] = {0.0, 0.0, 0.0};
};
struct body_dyn_t // what build_contact_jacobian reads of an object
{
    v3 pos, vel, ang_vel;
    quat q;
    double inv_mass;
    m3 inv_inertia_world;
    double restitution, friction;
};

// constraint.h:104-113
inline std::pair<v3, v3> build_orthonormal_basis(v3 n)
{
    double sign = std::copysign(1.0, n.z);
    const double a = -1.0 / (sign + n.z);
    const double b = n.x * n.y * a;
    return {v3{1.0 + sign * n.x * n.x * a, sign * b, -sign * n.x}, v3{b, sign + n.y * n.y * a, -n.y}};
}

// constraint.h:874-953 followed by the per-contact part of setup_contacts (:1069-1099)
inline std::optional<contact_solver_point_t> setup_contact(const body_dyn_t &a, const body_dyn_t &b, const contact_info_t &contact,
                                                           double dt, double gravity_norm)
{
    constexpr double bias_factor = .1;
    constexpr double linear_slop = 0.005;
    const double restitution_threshold = 2 * gravity_norm * dt; // :1071
    v3 nn = normalized(contact.normal);
    v3 local_normal_b = rotate(conjugate(b.q), nn);
    v3 r_a = rotate(a.q, contact.local_a);
    v3 r_b = rotate(b.q, contact.local_b);
    v3 n = rotate(b.q, local_normal_b);
    double penetration = dot((b.pos + r_b) - (a.pos + r_a), n);
    if (penetration <= 0.0) return std::nullopt;

    contact_solver_point_t p;
    auto m_eff = [&](const jacobian_row_t &row)
    {
        v3 k_a = mul(a.inv_inertia_world, row.J_w_a);
        v3 k_b = mul(b.inv_inertia_world, row.J_w_b);
        return 1.0 / (a.inv_mass + b.inv_mass + dot(row.J_w_a, k_a) + dot(row.J_w_b, k_b));
    };
    {
        jacobian_row_t row;
        row.J_v = n;
        row.J_w_a = cross(r_a, n);
        row.J_w_b = -cross(r_b, n);
        row.M_eff = m_eff(row);
        v3 v_ca = a.vel + cross(a.ang_vel, r_a);
        v3 v_cb = b.vel + cross(b.ang_vel, r_b);
        double v_rel_n = dot(v_ca - v_cb, n);
        double restitution_coeff = std::max(a.restitution, b.restitution);
        double restitution_bias = 0.0;
        if (v_rel_n < -restitution_threshold) restitution_bias = restitution_coeff * v_rel_n;
        double baumgarte_bias = (-bias_factor / dt) * std::max(0.0, penetration - linear_slop);
        row.bias = std::min(baumgarte_bias, restitution_bias);
        p.normal = row;
    }
    auto build_tangent = [&](v3 tangent)
    {
        jacobian_row_t row;
        row.J_v = tangent;
        row.J_w_a = cross(r_a, tangent);
        row.J_w_b = -cross(r_b, tangent);
        row.bias = 0.0;
        row.M_eff = m_eff(row);
        return row;
    };
    auto [t1, t2] = build_orthonormal_basis(n);
    p.tangent1 = build_tangent(t1);
    p.tangent2 = build_tangent(t2);

    p.friction_coeff = std::sqrt(a.friction * b.friction);
    double m11 = 1.0 / p.tangent1.M_eff;
    double m22 = 1.0 / p.tangent2.M_eff;
    double m12 = dot(p.tangent1.J_w_a, mul(a.inv_inertia_world, p.tangent2.J_w_a)) +
                 dot(p.tangent1.J_w_b, mul(b.inv_inertia_world, p.tangent2.J_w_b));
    double det = m11 * m22 - m12 * m12;
    if (det > 0.0)
    {
        p.inv_m_11 = m22 / det;
        p.inv_m_22 = m11 / det;
        p.inv_m_12 = -m12 / det;
    }
    else
    {
        p.inv_m_11 = p.tangent1.M_eff;
        p.inv_m_22 = p.tangent2.M_eff;
        p.inv_m_12 = 0.0;
    }
    p.accumulated[0] = contact.normal_impulse;
    p.accumulated[1] = contact.tangent_impulses[0];
    p.accumulated[2] = contact.tangent_impulses[1];
    return p;
}

// collision.cpp:512-518
inline std::optional<collision_info> gjk_epa(const shape &a, const shape &b, gjk_stats *st = nullptr)
{
    auto s = gjk_collision(a, b, st);
    if (s) return epa_solver::solve(a, b, *s, st);
    return std::nullopt;
}

// ---------------------------------------------------------------------------------------------
// Closest distance between two convex bodies by brute force — the checker of pk_gjk_distance_batch.
// The reference has no distance query (gjk_collision is boolean, src/collision.cpp:165-189), so there is nothing to
// restate; this is an independent method that shares no step with GJK: a polytope is the hull of its vertices
// (box: its 8 corners through the same rotate() the support mapping uses, hull: mesh vertices, sphere: its centre
// with a margin of r), and the closest pair of two disjoint polytopes is vertex–face, edge–edge or a degenerate
// form of those.  Every face is covered by triangles of vertex triples and every edge is a vertex pair, and any
// triple or pair spans points OF the hull, so
//     min( point–triangle over all vertices × all triples of the other body, both ways;
//          segment–segment over all vertex pairs × all vertex pairs )
// is the distance (O(n⁴): for the ≤ 16-vertex bodies of the tests).  Meaningful for disjoint bodies only.
// ---------------------------------------------------------------------------------------------
inline double brute_point_segment_sq(v3 p, v3 a, v3 b)
{
    const v3 ab = b - a;
    const double den = sqnorm(ab);
    double t = den > 0.0 ? dot(p - a, ab) / den : 0.0;
    t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
    return sqnorm(p - (a + t * ab));
}
inline double brute_point_triangle_sq(v3 p, v3 a, v3 b, v3 c)
{
    // foot of the perpendicular in the triangle's plane by its 2×2 normal equations; inside → that, else the border
    const v3 e0 = b - a, e1 = c - a, d = p - a;
    const double a00 = dot(e0, e0), a01 = dot(e0, e1), a11 = dot(e1, e1), b0 = dot(d, e0), b1 = dot(d, e1);
    const double det = a00 * a11 - a01 * a01;
    double best = std::min({brute_point_segment_sq(p, a, b), brute_point_segment_sq(p, b, c), brute_point_segment_sq(p, c, a)});
    if (det > 1e-14 * a00 * a11)
    {
        const double u = (a11 * b0 - a01 * b1) / det, w = (a00 * b1 - a01 * b0) / det;
        if (u >= 0.0 && w >= 0.0 && u + w <= 1.0) best = std::min(best, sqnorm(d - (u * e0 + w * e1)));
    }
    return best;
}
inline double brute_segment_segment_sq(v3 p1, v3 q1, v3 p2, v3 q2)
{
    // a convex quadratic over the unit square: the stationary point if it lies inside, else the border
    const v3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
    const double a = dot(d1, d1), e = dot(d2, d2), b = dot(d1, d2), c = dot(d1, r), f = dot(d2, r);
    double best = std::min({brute_point_segment_sq(p1, p2, q2), brute_point_segment_sq(q1, p2, q2), brute_point_segment_sq(p2, p1, q1),
                            brute_point_segment_sq(q2, p1, q1)});
    const double den = a * e - b * b;
    if (den > 1e-14 * a * e)
    {
        const double s = (b * f - c * e) / den, t = (a * f - b * c) / den;
        if (s >= 0.0 && s <= 1.0 && t >= 0.0 && t <= 1.0) best = std::min(best, sqnorm((p1 + s * d1) - (p2 + t * d2)));
    }
    return best;
}
// world-frame vertices of a body's core and its margin
inline double brute_core(const shape &s, std::vector<v3> &out)
{
    out.clear();
    switch (s.kind)
    {
    case KIND_SPHERE:
        out.push_back(s.a);
        return s.b.x;
    case KIND_AABB:
        for (int i = 0; i < 8; ++i) out.push_back({(i & 1) ? s.b.x : s.a.x, (i & 2) ? s.b.y : s.a.y, (i & 4) ? s.b.z : s.a.z});
        return 0.0;
    case KIND_OBB:
        for (int i = 0; i < 8; ++i)
            out.push_back(s.a + rotate(s.q, v3{(i & 1) ? s.b.x : -s.b.x, (i & 2) ? s.b.y : -s.b.y, (i & 4) ? s.b.z : -s.b.z}));
        return 0.0;
    default:
        for (uint32_t i = 0; i < s.nverts; ++i) out.push_back(rotate(s.q, s.verts[i]) + s.a);
        return 0.0;
    }
}
// → distance of the two bodies if they are disjoint (negative or zero: the margins overlap); NaN above max_verts
inline double brute_distance(const shape &sa, const shape &sb, uint32_t max_verts = 16)
{
    std::vector<v3> A, B;
    const double ra = brute_core(sa, A), rb = brute_core(sb, B);
    if (A.size() > max_verts || B.size() > max_verts) return std::numeric_limits<double>::quiet_NaN();
    double best = std::numeric_limits<double>::infinity();
    auto point_vs = [&](const std::vector<v3> &P, const std::vector<v3> &Q)
    {
        const std::size_t n = Q.size();
        for (const v3 &p : P)
        {
            if (n == 1) best = std::min(best, sqnorm(p - Q[0]));
            for (std::size_t i = 0; i < n; ++i)
                for (std::size_t j = i + 1; j < n; ++j)
                {
                    if (n == 2) best = std::min(best, brute_point_segment_sq(p, Q[i], Q[j]));
                    for (std::size_t k = j + 1; k < n; ++k) best = std::min(best, brute_point_triangle_sq(p, Q[i], Q[j], Q[k]));
                }
        }
    };
    point_vs(A, B);
    point_vs(B, A);
    for (std::size_t i = 0; i < A.size(); ++i)
        for (std::size_t j = i + 1; j < A.size(); ++j)
            for (std::size_t k = 0; k < B.size(); ++k)
                for (std::size_t l = k + 1; l < B.size(); ++l) best = std::min(best, brute_segment_segment_sq(A[i], A[j], B[k], B[l]));
    return std::sqrt(best) - ra - rb;
}

// ---------------------------------------------------------------------------------------------
// dynamic_bvh — include/physkit/collision/bvh.h:270-535, src/bvh.cpp:239-514
// ---------------------------------------------------------------------------------------------
class dynamic_bvh
{
public:
    static constexpr std::size_t stack_size = 64; // bvh.h:273
    static constexpr uint32_t null = std::numeric_limits<uint32_t>::max();
    static constexpr uint32_t free_node_h = std::numeric_limits<uint32_t>::max();

    struct node // bvh.h:465-501
    {
        aabb bounds{};
        uint32_t parent = null;
        union
        {
            struct
            {
                uint32_t left, right;
            } children;
            uint32_t data;
            uint32_t next_free;
        };
        uint32_t height = free_node_h;
        node() { children.left = null; children.right = null; }
        bool is_leaf() const { return height == 0; }
        bool is_free() const { return height == free_node_h; }
    };

    explicit dynamic_bvh(std::size_t cap = 1024) { M_nodes.reserve(cap); }

    // bvh.h:294-302
    uint32_t add(uint32_t id, const aabb &bounds)
    {
        uint32_t leaf = allocate_node();
        M_nodes[leaf].data = id;
        M_nodes[leaf].bounds = bounds;
        insert_leaf(leaf);
        refit_and_rotate(M_nodes[leaf].parent);
        return leaf;
    }

    // bvh.h:307-311
    void remove_leaf(uint32_t leaf)
    {
        extract_leaf(leaf);
        free_node(leaf);
    }

    // src/bvh.cpp:475-514
    bool update_leaf(uint32_t leaf, const aabb &true_bounds, v3 disp)
    {
        assert(!M_nodes[leaf].is_free());
        assert(M_nodes[leaf].is_leaf());
        if (M_nodes[leaf].bounds.contains(true_bounds)) return false;
        extract_leaf(leaf);
        aabb nb = M_nodes[leaf].bounds;
        // fat_update re-tests contains() (already known false) and applies the margin rule.
        fat_update(nb, true_bounds, disp);
        M_nodes[leaf].bounds = nb;
        insert_leaf(leaf);
        refit_and_rotate(M_nodes[leaf].parent);
        return true;
    }

    // bvh.h:313-344
    template <typename F> void query_aabb(const aabb &box, F &&callback) const
    {
        if (M_root == null) return;
        std::array<uint32_t, stack_size> stack{};
        int sp = 0;
        stack[sp++] = M_root;
        while (sp > 0)
        {
            uint32_t idx = stack[--sp];
            const node &n = M_nodes[idx];
            if (!n.bounds.intersects(box)) continue;
            if (n.is_leaf())
            {
                if (!callback(n.data)) return;
                continue;
            }
            assert(sp + 1 < static_cast<int>(stack_size));
            bool left_first = M_nodes[n.children.left].bounds.surface_area() >
                              M_nodes[n.children.right].bounds.surface_area();
            uint32_t first = left_first ? n.children.left : n.children.right;
            uint32_t second = left_first ? n.children.right : n.children.left;
            stack[sp++] = first;
            stack[sp++] = second;
        }
    }

    // bvh.h:346-398 — callback form: callback(data, d_node, max_distance) returns the new max distance,
    // ≤ 0 ends the cast.  Nearer child visited first.
    template <typename F> void raycast(const ray &r, double max_distance, F &&callback) const
    {
        if (M_root == null) return;
        std::array<uint32_t, stack_size> stack{};
        int sp = 0;
        stack[sp++] = M_root;
        while (sp > 0)
        {
            uint32_t idx = stack[--sp];
            const node &n = M_nodes[idx];
            auto d_node = r.intersect_distance(n.bounds, max_distance);
            if (!d_node.has_value()) continue;
            if (n.is_leaf())
            {
                max_distance = callback(n.data, *d_node, max_distance);
                if (max_distance <= 0.0) return;
                continue;
            }
            assert(sp + 1 < static_cast<int>(stack_size));
            uint32_t left = n.children.left, right = n.children.right;
            auto d_left = r.intersect_distance(M_nodes[left].bounds, max_distance);
            auto d_right = r.intersect_distance(M_nodes[right].bounds, max_distance);
            if (d_left && d_right)
            {
                if (*d_left > *d_right)
                {
                    stack[sp++] = left;
                    stack[sp++] = right;
                }
                else
                {
                    stack[sp++] = right;
                    stack[sp++] = left;
                }
            }
            else if (d_left)
                stack[sp++] = left;
            else if (d_right)
                stack[sp++] = right;
        }
    }

    // bvh.h:400-450 — generator form: every leaf whose box the ray enters within max_distance, in the
    // order the coroutine yields them (same traversal, max_distance fixed).
    std::vector<std::pair<uint32_t, double>> raycast(const ray &r, double max_distance) const
    {
        std::vector<std::pair<uint32_t, double>> out;
        raycast(r, max_distance,
                [&](uint32_t data, double d, double md)
                {
                    out.emplace_back(data, d);
                    return md > 0.0 ? md : std::numeric_limits<double>::min(); // the generator never stops early
                });
        return out;
    }

    const aabb &bounds(uint32_t leaf) const { return M_nodes[leaf].bounds; }
    uint32_t data(uint32_t leaf) const { return M_nodes[leaf].data; }
    uint32_t root() const { return M_root; }
    const std::vector<node> &nodes() const { return M_nodes; }

    // structural invariant check used by tests (not part of the reference API)
    bool validate() const
    {
        if (M_root == null) return true;
        return validate_node(M_root, null) >= 0;
    }

private:
    std::vector<node> M_nodes;
    uint32_t M_root = null;
    uint32_t M_free_head = null;

    int validate_node(uint32_t idx, uint32_t parent) const
    {
        const node &n = M_nodes[idx];
        if (n.parent != parent || n.is_free()) return -1;
        if (n.is_leaf()) return 0;
        int hl = validate_node(n.children.left, idx);
        int hr = validate_node(n.children.right, idx);
        if (hl < 0 || hr < 0) return -1;
        aabb u = aabb_union(M_nodes[n.children.left].bounds, M_nodes[n.children.right].bounds);
        if (std::memcmp(&u, &n.bounds, sizeof(aabb)) != 0) return -1;
        int h = 1 + std::max(hl, hr);
        if (static_cast<uint32_t>(h) != n.height) return -1;
        return h;
    }

    // bvh.h:507-529
    uint32_t allocate_node()
    {
        if (M_free_head != null)
        {
            uint32_t a = M_free_head;
            M_free_head = M_nodes[a].next_free;
            M_nodes[a].parent = null;
            M_nodes[a].height = 0;
            return a;
        }
        M_nodes.emplace_back().height = 0;
        return static_cast<uint32_t>(M_nodes.size() - 1);
    }
    void free_node(uint32_t n)
    {
        M_nodes[n].height = free_node_h;
        M_nodes[n].next_free = M_free_head;
        M_free_head = n;
    }

    // src/bvh.cpp:239-318
    uint32_t insert_leaf(uint32_t leaf_idx)
    {
        constexpr double cost_traversal = 2.0;
        constexpr double cost_make = 2.0;
        if (M_root == null)
        {
            M_root = leaf_idx;
            M_nodes[leaf_idx].parent = null;
            return leaf_idx;
        }
        aabb leaf_bounds = M_nodes[leaf_idx].bounds;
        uint32_t cur = M_root;
        while (!M_nodes[cur].is_leaf())
        {
            uint32_t left = M_nodes[cur].children.left;
            uint32_t right = M_nodes[cur].children.right;
            double area_cur = M_nodes[cur].bounds.surface_area();
            aabb nb = aabb_union(M_nodes[cur].bounds, leaf_bounds);
            double area_new = nb.surface_area();
            double cost_new_sibling = cost_make * area_new;
            double inherited = cost_traversal * (area_new - area_cur);
            double cost_left, cost_right;
            if (M_nodes[left].is_leaf())
                cost_left = cost_traversal * aabb_union(leaf_bounds, M_nodes[left].bounds).surface_area() + inherited;
            else
                cost_left = cost_traversal * (aabb_union(leaf_bounds, M_nodes[left].bounds).surface_area() -
                                              M_nodes[left].bounds.surface_area()) + inherited;
            if (M_nodes[right].is_leaf())
                cost_right = cost_traversal * aabb_union(leaf_bounds, M_nodes[right].bounds).surface_area() + inherited;
            else
                cost_right = cost_traversal * (aabb_union(leaf_bounds, M_nodes[right].bounds).surface_area() -
                                               M_nodes[right].bounds.surface_area()) + inherited;
            if (cost_new_sibling < cost_left && cost_new_sibling < cost_right) break;
            cur = (cost_left < cost_right) ? left : right;
        }
        uint32_t sibling = cur;
        uint32_t old_parent = M_nodes[sibling].parent;
        uint32_t new_parent = allocate_node();
        node &np = M_nodes[new_parent];
        np.bounds = aabb_union(leaf_bounds, M_nodes[sibling].bounds);
        np.parent = old_parent;
        np.children.left = sibling;
        np.children.right = leaf_idx;
        np.height = M_nodes[sibling].height + 1;
        M_nodes[sibling].parent = new_parent;
        M_nodes[leaf_idx].parent = new_parent;
        if (old_parent != null)
        {
            if (M_nodes[old_parent].children.left == sibling)
                M_nodes[old_parent].children.left = new_parent;
            else
                M_nodes[old_parent].children.right = new_parent;
        }
        else
            M_root = new_parent;
        return leaf_idx;
    }

    // src/bvh.cpp:320-355
    void extract_leaf(uint32_t leaf_idx)
    {
        if (leaf_idx == M_root)
        {
            M_root = null;
            return;
        }
        uint32_t parent = M_nodes[leaf_idx].parent;
        uint32_t grand = M_nodes[parent].parent;
        uint32_t sibling = (M_nodes[parent].children.left == leaf_idx) ? M_nodes[parent].children.right
                                                                       : M_nodes[parent].children.left;
        if (grand != null)
        {
            if (M_nodes[grand].children.left == parent)
                M_nodes[grand].children.left = sibling;
            else
                M_nodes[grand].children.right = sibling;
            M_nodes[sibling].parent = grand;
            free_node(parent);
            refit_and_rotate(grand);
        }
        else
        {
            M_root = sibling;
            M_nodes[sibling].parent = null;
            free_node(parent);
        }
        M_nodes[leaf_idx].parent = null;
    }

    // src/bvh.cpp:357-372
    void refit_and_rotate(uint32_t idx)
    {
        while (idx != null)
        {
            idx = balance(idx);
            node &n = M_nodes[idx];
            node &l = M_nodes[n.children.left];
            node &r = M_nodes[n.children.right];
            n.height = 1 + std::max(l.height, r.height);
            n.bounds = aabb_union(l.bounds, r.bounds);
            idx = n.parent;
        }
    }

    // src/bvh.cpp:374-473
    uint32_t balance(uint32_t i_a)
    {
        node &a = M_nodes[i_a];
        if (a.is_leaf() || a.height < 2) return i_a;
        uint32_t i_b = a.children.left, i_c = a.children.right;
        node &b = M_nodes[i_b];
        node &c = M_nodes[i_c];
        int bal = static_cast<int>(c.height) - static_cast<int>(b.height);
        if (bal > 1)
        {
            uint32_t i_f = c.children.left, i_g = c.children.right;
            node &f = M_nodes[i_f];
            node &g = M_nodes[i_g];
            c.children.left = i_a;
            c.parent = a.parent;
            a.parent = i_c;
            if (c.parent != null)
            {
                if (M_nodes[c.parent].children.left == i_a)
                    M_nodes[c.parent].children.left = i_c;
                else
                    M_nodes[c.parent].children.right = i_c;
            }
            else
                M_root = i_c;
            if (f.height > g.height)
            {
                c.children.right = i_f;
                a.children.right = i_g;
                g.parent = i_a;
            }
            else
            {
                c.children.right = i_g;
                a.children.right = i_f;
                f.parent = i_a;
            }
            a.bounds = aabb_union(b.bounds, M_nodes[a.children.right].bounds);
            a.height = 1 + std::max(M_nodes[a.children.left].height, M_nodes[a.children.right].height);
            return i_c;
        }
        if (bal < -1)
        {
            uint32_t i_d = b.children.left, i_e = b.children.right;
            node &d = M_nodes[i_d];
            node &e = M_nodes[i_e];
            b.children.left = i_a;
            b.parent = a.parent;
            a.parent = i_b;
            if (b.parent != null)
            {
                if (M_nodes[b.parent].children.left == i_a)
                    M_nodes[b.parent].children.left = i_b;
                else
                    M_nodes[b.parent].children.right = i_b;
            }
            else
                M_root = i_b;
            if (d.height > e.height)
            {
                b.children.right = i_d;
                a.children.left = i_e;
                e.parent = i_a;
            }
            else
            {
                b.children.right = i_e;
                a.children.left = i_d;
                d.parent = i_a;
            }
            a.bounds = aabb_union(c.bounds, M_nodes[a.children.left].bounds);
            a.height = 1 + std::max(M_nodes[a.children.left].height, M_nodes[a.children.right].height);
            return i_b;
        }
        return i_a;
    }
};

// ---------------------------------------------------------------------------------------------
// pair_manager + broad_phase — include/physkit/collision/collision_phases.h:29-73, 330-445.
// The active set is an unordered set of u64 keys; only its CONTENT is contractually observable
// (abseil iteration order affects manifold order, never values), so sorted_pairs() is the result.
// ---------------------------------------------------------------------------------------------
inline uint64_t make_pair_key(uint32_t a, uint32_t b) // collision_phases.h:63-69
{
    uint32_t mn = std::min(a, b), mx = std::max(a, b);
    return (static_cast<uint64_t>(mn) << 32) | mx;
}

class broad_phase
{
public:
    struct handle
    {
        uint32_t node;
        bool is_static;
    };

    // collision_phases.h:342-346
    handle add(uint32_t object_id, const aabb &bounds, bool is_static)
    {
        uint32_t n = is_static ? M_static.add(object_id, bounds) : M_dynamic.add(object_id, bounds);
        if (object_id >= M_handles.size()) M_handles.resize(object_id + 1, handle{dynamic_bvh::null, false});
        M_handles[object_id] = {n, is_static};
        return M_handles[object_id];
    }

    // collision_phases.h:350-369
    void remove(uint32_t object_id)
    {
        handle h = M_handles[object_id];
        if (h.is_static)
            M_static.remove_leaf(h.node);
        else
        {
            M_dynamic.remove_leaf(h.node);
            M_moved.erase(std::remove(M_moved.begin(), M_moved.end(), h.node), M_moved.end());
        }
        for (auto it = M_active.begin(); it != M_active.end();)
        {
            uint32_t a = static_cast<uint32_t>(*it >> 32), b = static_cast<uint32_t>(*it & 0xFFFFFFFFu);
            if (a == object_id || b == object_id)
                it = M_active.erase(it);
            else
                ++it;
        }
        M_handles[object_id] = {dynamic_bvh::null, false};
    }

    // collision_phases.h:371-375
    bool update_node(uint32_t object_id, const aabb &bounds, v3 disp)
    {
        handle h = M_handles[object_id];
        assert(!h.is_static);
        if (M_dynamic.update_leaf(h.node, bounds, disp))
        {
            M_moved.push_back(h.node);
            return true;
        }
        return false;
    }

    // collision_phases.h:377-436
    void calculate_pairs()
    {
        for (auto it = M_active.begin(); it != M_active.end();)
        {
            uint64_t key = *it;
            uint32_t a = static_cast<uint32_t>(key >> 32), b = static_cast<uint32_t>(key & 0xFFFFFFFFu);
            const aabb &ba = stored(a);
            const aabb &bb = stored(b);
            if (!ba.intersects(bb))
                it = M_active.erase(it);
            else
                ++it;
        }
        for (uint32_t id : M_moved)
        {
            const aabb moving = M_dynamic.bounds(id);
            const uint32_t ea = M_dynamic.data(id);
            auto cb = [&](uint32_t eb)
            {
                if (ea != eb) M_active.insert(make_pair_key(ea, eb));
                return true;
            };
            M_dynamic.query_aabb(moving, cb);
            M_static.query_aabb(moving, cb);
        }
        M_moved.clear();
    }

    const aabb &stored(uint32_t object_id) const
    {
        handle h = M_handles[object_id];
        return h.is_static ? M_static.bounds(h.node) : M_dynamic.bounds(h.node);
    }

    // world_base::raycast (core/world.h:260-319): the static and the dynamic tree's streams merged by
    // distance, the static entry first on equal distances.
    std::vector<std::pair<uint32_t, double>> raycast(const ray &r, double max_dist) const
    {
        auto st = M_static.raycast(r, max_dist);
        auto dy = M_dynamic.raycast(r, max_dist);
        std::vector<std::pair<uint32_t, double>> out;
        std::size_t i = 0, j = 0;
        while (i < st.size() && j < dy.size())
        {
            if (st[i].second <= dy[j].second)
                out.push_back(st[i++]);
            else
                out.push_back(dy[j++]);
        }
        while (i < st.size()) out.push_back(st[i++]);
        while (j < dy.size()) out.push_back(dy[j++]);
        return out;
    }

    std::vector<uint64_t> sorted_pairs() const
    {
        std::vector<uint64_t> v(M_active.begin(), M_active.end());
        std::sort(v.begin(), v.end());
        return v;
    }
    std::size_t moved_count() const { return M_moved.size(); }
    const dynamic_bvh &dynamic_tree() const { return M_dynamic; }
    const dynamic_bvh &static_tree() const { return M_static; }

private:
    dynamic_bvh M_static;
    dynamic_bvh M_dynamic;
    std::vector<uint32_t> M_moved;
    std::unordered_set<uint64_t> M_active;
    std::vector<handle> M_handles;
};

} // namespace pko
