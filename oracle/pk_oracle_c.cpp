// pk_oracle_c.cpp — C ABI over pk_oracle.hpp for ctypes.  CPU ORACLE: test infrastructure only
// (see the header of pk_oracle.hpp).  Nothing in physkit_b200/ may link or load this.
#include "pk_oracle.hpp"

#include <map>

#include <atomic>
#include <cstdio>
#include <thread>

using namespace pko;

extern "C"
{

// Mirrors pk_shape_* of include/pk_collide.h (the product ABI) so the same scene arrays can be fed
// to both sides.
struct pko_shape_desc
{
    int32_t kind;      // shape_kind
    uint32_t vert_off; // HULL: first vertex (in units of vertices) in the shared vertex pool
    uint32_t nverts;   // HULL
    uint32_t _pad;
    double a[3];    // AABB: min | OBB: half extents | SPHERE: a[0] = radius
    double b[3];    // AABB: max
    double lmin[3]; // local AABB (what mesh::bounds() would be, mesh.h:191,248)
    double lmax[3];
};

static inline v3 ld3(const double *p) { return {p[0], p[1], p[2]}; }
static inline quat ldq(const double *p) { return {p[0], p[1], p[2], p[3]}; } // x,y,z,w
static inline void st3(double *o, v3 v)
{
    o[0] = v.x;
    o[1] = v.y;
    o[2] = v.z;
}

// Fill lmin/lmax from the shape parameters.
void pko_shape_local_aabb(pko_shape_desc *s, const double *verts)
{
    switch (s->kind)
    {
    case KIND_AABB:
        for (int k = 0; k < 3; ++k) s->lmin[k] = s->a[k], s->lmax[k] = s->b[k];
        break;
    case KIND_OBB:
        for (int k = 0; k < 3; ++k) s->lmin[k] = -s->a[k], s->lmax[k] = s->a[k];
        break;
    case KIND_SPHERE:
        for (int k = 0; k < 3; ++k) s->lmin[k] = -s->a[0], s->lmax[k] = s->a[0];
        break;
    default:
    {
        aabb bx = aabb_from_points(reinterpret_cast<const v3 *>(verts) + s->vert_off, s->nverts);
        st3(s->lmin, bx.min);
        st3(s->lmax, bx.max);
    }
    }
}

static inline aabb body_bounds(const pko_shape_desc &s, const double *pos, const double *q)
{
    aabb local{ld3(s.lmin), ld3(s.lmax)};
    if (s.kind == KIND_AABB) return local; // pose-less shape (gjk_epa(aabb, …) instantiations only)
    return instance_bounds(local, ld3(pos), ldq(q));
}

static inline shape make_shape(const pko_shape_desc &s, const double *verts, const double *pos,
                               const double *q)
{
    shape sh{};
    sh.kind = s.kind;
    switch (s.kind)
    {
    case KIND_AABB:
        sh.a = ld3(s.a);
        sh.b = ld3(s.b);
        break;
    case KIND_OBB:
        sh.a = ld3(pos);
        sh.b = ld3(s.a);
        sh.q = ldq(q);
        break;
    case KIND_SPHERE:
        sh.a = ld3(pos);
        sh.b = v3{s.a[0], 0, 0};
        break;
    default:
        sh.a = ld3(pos);
        sh.q = ldq(q);
        sh.verts = reinterpret_cast<const v3 *>(verts) + s.vert_off;
        sh.nverts = s.nverts;
    }
    return sh;
}

// mesh::instance::bounds() for every body → out6[n] = (min xyz, max xyz).
void pko_bounds(const pko_shape_desc *shapes, const double *pos, const double *quat_xyzw,
                const uint32_t *shape_id, uint64_t n, double *out6)
{
    for (uint64_t i = 0; i < n; ++i)
    {
        aabb b = body_bounds(shapes[shape_id[i]], pos + 3 * i, quat_xyzw + 4 * i);
        st3(out6 + 6 * i, b.min);
        st3(out6 + 6 * i + 3, b.max);
    }
}

// Single support query (tests/obb/obb_test.cpp:142-159, tests/mesh/main.cpp:890-899,1660-1710).
void pko_support(const pko_shape_desc *shapes, const double *verts, const double *pos,
                 const double *quat_xyzw, uint32_t shape_id, const double *dir, double *out3)
{
    shape s = make_shape(shapes[shape_id], verts, pos, quat_xyzw);
    st3(out3, support(s, ld3(dir)));
}

// brute_distance(body pair_a[k], body pair_b[k]) for every k (checker of pk_gjk_distance_batch; NaN for bodies with more
// than max_verts vertices).  Independent pairs over std::thread workers.
void pko_distance_brute_pairs(const pko_shape_desc *shapes, const double *verts, const double *pos, const double *quat_xyzw,
                              const uint32_t *shape_id, const uint32_t *pair_a, const uint32_t *pair_b, uint64_t npairs,
                              uint32_t max_verts, double *out, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    std::atomic<uint64_t> next{0};
    auto worker = [&]()
    {
        for (;;)
        {
            const uint64_t k0 = next.fetch_add(64);
            if (k0 >= npairs) break;
            const uint64_t k1 = std::min<uint64_t>(npairs, k0 + 64);
            for (uint64_t k = k0; k < k1; ++k)
            {
                const uint32_t ia = pair_a[k], ib = pair_b[k];
                shape a = make_shape(shapes[shape_id[ia]], verts, pos + 3 * ia, quat_xyzw + 4 * ia);
                shape b = make_shape(shapes[shape_id[ib]], verts, pos + 3 * ib, quat_xyzw + 4 * ib);
                out[k] = brute_distance(a, b, max_verts);
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
}

// gjk_epa(a = body pair_a[k], b = body pair_b[k]) for every k.
//   out10[k] = normal(3) world_a(3) world_b(3) depth ; hit[k] ∈ {0,1} ; stats8[k] optional.
// nthreads > 1 uses std::thread workers over independent pairs (courtesy all-cores figure; the reference
// itself is single-threaded: collision_phases.h:251 "TODO: parallelize").
uint64_t pko_gjk_epa_pairs(const pko_shape_desc *shapes, const double *verts, const double *pos,
                           const double *quat_xyzw, const uint32_t *shape_id, const uint32_t *pair_a,
                           const uint32_t *pair_b, uint64_t npairs, double *out10, uint8_t *hit,
                           int32_t *stats8, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    std::atomic<uint64_t> next{0};
    std::atomic<uint64_t> nhit{0};
    constexpr uint64_t chunk = 256;
    auto worker = [&]()
    {
        uint64_t local_hits = 0;
        for (;;)
        {
            uint64_t k0 = next.fetch_add(chunk);
            if (k0 >= npairs) break;
            uint64_t k1 = std::min(npairs, k0 + chunk);
            for (uint64_t k = k0; k < k1; ++k)
            {
                uint32_t ia = pair_a[k], ib = pair_b[k];
                shape a = make_shape(shapes[shape_id[ia]], verts, pos + 3 * ia, quat_xyzw + 4 * ia);
                shape b = make_shape(shapes[shape_id[ib]], verts, pos + 3 * ib, quat_xyzw + 4 * ib);
                gjk_stats st;
                auto r = gjk_epa(a, b, &st);
                if (hit) hit[k] = r ? 1 : 0;
                if (out10)
                {
                    double *o = out10 + 10 * k;
                    if (r)
                    {
                        st3(o, r->normal);
                        st3(o + 3, r->world_a);
                        st3(o + 6, r->world_b);
                        o[9] = r->depth;
                    }
                    else
                        for (int j = 0; j < 10; ++j) o[j] = 0.0;
                }
                if (stats8)
                {
                    int32_t *s = stats8 + 8 * k;
                    s[0] = st.gjk_iters;
                    s[1] = st.epa_iters;
                    s[2] = st.epa_faces;
                    s[3] = st.epa_verts;
                    s[4] = st.epa_heap_max;
                    s[5] = st.epa_horizon_max;
                    s[6] = st.epa_stack_max;
                    s[7] = st.exit_code;
                }
                if (r) ++local_hits;
            }
        }
        nhit += local_hits;
    };
    if (nthreads == 1)
        worker();
    else
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) pool.emplace_back(worker);
        for (auto &t : pool) t.join();
    }
    return nhit.load();
}

// contact_point(info, a, b) for every contact (collision_phases.h:78-82): the witness points in the bodies'
// own frames, particle::project_to_local = orientation.conjugate() * (world_point - pos) (core/particle.h:107-108).
//   pair_a/pair_b: bodies of contact k; contacts10[k] = normal(3) world_a(3) world_b(3) depth; out6[k] = local_a local_b
void pko_contact_points(const double *pos, const double *quat_xyzw, const uint32_t *pair_a, const uint32_t *pair_b,
                        const double *contacts10, uint64_t n, double *out6)
{
    for (uint64_t k = 0; k < n; ++k)
    {
        const uint32_t ia = pair_a[k], ib = pair_b[k];
        const quat qa{quat_xyzw[4 * ia], quat_xyzw[4 * ia + 1], quat_xyzw[4 * ia + 2], quat_xyzw[4 * ia + 3]};
        const quat qb{quat_xyzw[4 * ib], quat_xyzw[4 * ib + 1], quat_xyzw[4 * ib + 2], quat_xyzw[4 * ib + 3]};
        const double *c = contacts10 + 10 * k;
        st3(out6 + 6 * k, rotate(conjugate(qa), ld3(c + 3) - ld3(pos + 3 * ia)));
        st3(out6 + 6 * k + 3, rotate(conjugate(qb), ld3(c + 6) - ld3(pos + 3 * ib)));
    }
}

// ------------------------------- manifolds (narrow_phase state) --------------------------------
// narrow_phase keeps one manifold per pair of the pair set (collision_phases.h:218-242, 322-327); only the
// non-empty ones carry state, so that is what is stored here, keyed by make_pair_key.
struct pko_manifolds
{
    std::map<uint64_t, manifold_t> mans;
};
void *pko_man_create() { return new pko_manifolds(); }
void pko_man_destroy(void *h) { delete static_cast<pko_manifolds *>(h); }

// narrow_phase::calculate (collision_phases.h:244-320) for the current pair set.
//   keys[n] sorted pair keys, hit[n], contacts10[n] (rows of hits: normal, world_a, world_b, depth; others ignored)
//   began / ended: keys whose manifold went empty → non-empty / non-empty → empty this step (on_coll_beg / on_coll_end)
// Pairs that left the pair set are dropped without a callback (on_pair_removed, :225-242).  Returns the number
// of non-empty manifolds.
uint64_t pko_man_step(void *h, const uint64_t *keys, const uint8_t *hit, const double *contacts10, uint64_t n, const double *pos,
                      const double *quat_xyzw, uint64_t *began, uint64_t *nbegan, uint64_t *ended, uint64_t *nended)
{
    pko_manifolds &st = *static_cast<pko_manifolds *>(h);
    std::map<uint64_t, manifold_t> next;
    uint64_t nb = 0, ne = 0;
    for (uint64_t k = 0; k < n; ++k)
    {
        const uint64_t key = keys[k];
        const uint32_t ia = static_cast<uint32_t>(key >> 32), ib = static_cast<uint32_t>(key & 0xFFFFFFFFu);
        auto it = st.mans.find(key);
        if (it == st.mans.end() && !hit[k]) continue; // empty manifold, no new contact: nothing happens
        const manifold_t old_man = (it == st.mans.end()) ? manifold_t{} : it->second;
        const v3 pa = ld3(pos + 3 * ia), pb = ld3(pos + 3 * ib);
        const quat qa = ldq(quat_xyzw + 4 * ia), qb = ldq(quat_xyzw + 4 * ib);
        contact_info_t nc;
        if (hit[k])
        {
            const double *c = contacts10 + 10 * k;
            nc.normal = ld3(c);
            nc.local_a = rotate(conjugate(qa), ld3(c + 3) - pa); // contact_point (:78-82)
            nc.local_b = rotate(conjugate(qb), ld3(c + 6) - pb);
            nc.depth = c[9];
        }
        manifold_t nm = manifold_merge(old_man, hit[k] ? &nc : nullptr, pa, qa, pb, qb);
        const bool was = old_man.n > 0, is = nm.n > 0;
        if (!was && is && began) began[nb++] = key;
        if (was && !is && ended) ended[ne++] = key;
        if (is) next.emplace(key, nm);
    }
    st.mans.swap(next);
    if (nbegan) *nbegan = nb;
    if (nended) *nended = ne;
    return st.mans.size();
}
// keys_out[m], counts_out[m], pts[m][4][13] = normal(3) local_a(3) local_b(3) depth normal_impulse tangent(2), key order
uint64_t pko_man_get(void *h, uint64_t *keys_out, uint32_t *counts_out, double *pts, uint64_t cap)
{
    pko_manifolds &st = *static_cast<pko_manifolds *>(h);
    uint64_t m = 0;
    for (const auto &[key, man] : st.mans)
    {
        if (m >= cap) break;
        keys_out[m] = key;
        counts_out[m] = static_cast<uint32_t>(man.n);
        for (int j = 0; j < 4; ++j)
        {
            double *o = pts + (m * 4 + j) * 13;
            const contact_info_t &c = man.c[j];
            const bool used = j < man.n;
            st3(o, used ? c.normal : v3{0, 0, 0});
            st3(o + 3, used ? c.local_a : v3{0, 0, 0});
            st3(o + 6, used ? c.local_b : v3{0, 0, 0});
            o[9] = used ? c.depth : 0.0;
            o[10] = used ? c.normal_impulse : 0.0;
            o[11] = used ? c.tangent_impulses[0] : 0.0;
            o[12] = used ? c.tangent_impulses[1] : 0.0;
        }
        ++m;
    }
    return st.mans.size();
}
// what the constraint solver leaves behind (constraint.h:1107-1201 accumulates them): imp[m][4][3] in key order
void pko_man_set_impulses(void *h, const double *imp)
{
    pko_manifolds &st = *static_cast<pko_manifolds *>(h);
    uint64_t m = 0;
    for (auto &[key, man] : st.mans)
    {
        for (int j = 0; j < man.n; ++j)
        {
            man.c[j].normal_impulse = imp[(m * 4 + j) * 3];
            man.c[j].tangent_impulses[0] = imp[(m * 4 + j) * 3 + 1];
            man.c[j].tangent_impulses[1] = imp[(m * 4 + j) * 3 + 2];
        }
        ++m;
    }
}

// constraint_solver::setup_contacts (constraint.h:1052-1104) over the manifolds of h in key order, points in
// manifold order; bodies from flat arrays (inertia9: local tensor, row-major; the world inverse tensor is
// derived like particle::update_derived_state).  rows[k][44] = normal / tangent1 / tangent2 as
// (J_v 3, J_w_a 3, J_w_b 3, M_eff, bias) + friction_coeff, inv_m_11, inv_m_12, inv_m_22 + accumulated[3] = 40,
// padded to 44 with key (as two u32 halves stored in doubles would lose nothing, but keys_out carries it),
// point_out[k] = index of the point in its manifold.  Returns the number of rows (may exceed cap).
uint64_t pko_setup_contacts(void *h, const double *pos, const double *quat_xyzw, const double *vel, const double *ang_vel,
                            const double *mass, const double *inertia9, const double *restitution, const double *friction,
                            double dt, double gravity_norm, double *rows40, uint64_t *keys_out, uint32_t *point_out, uint64_t cap)
{
    pko_manifolds &st = *static_cast<pko_manifolds *>(h);
    auto body = [&](uint32_t i)
    {
        body_dyn_t b;
        b.pos = ld3(pos + 3 * i);
        b.vel = ld3(vel + 3 * i);
        b.ang_vel = ld3(ang_vel + 3 * i);
        b.q = ldq(quat_xyzw + 4 * i);
        b.inv_mass = 1.0 / mass[i];
        rigid_state o;
        o.q = b.q;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) o.inertia_local.m[r][c] = inertia9[9 * static_cast<uint64_t>(i) + 3 * r + c];
        o.inv_inertia_local = (b.inv_mass == 0.0) ? m3{} : inverse(o.inertia_local);
        o.update_derived_state();
        b.inv_inertia_world = o.inv_inertia_world;
        b.restitution = restitution[i];
        b.friction = friction[i];
        return b;
    };
    uint64_t k = 0;
    for (const auto &[key, man] : st.mans)
    {
        const uint32_t ia = static_cast<uint32_t>(key >> 32), ib = static_cast<uint32_t>(key & 0xFFFFFFFFu);
        const body_dyn_t a = body(ia), b = body(ib);
        for (int j = 0; j < man.n; ++j)
        {
            auto p = setup_contact(a, b, man.c[j], dt, gravity_norm);
            if (!p) continue;
            if (k < cap)
            {
                double *o = rows40 + 40 * k;
                const jacobian_row_t *rw[3] = {&p->normal, &p->tangent1, &p->tangent2};
                for (int r = 0; r < 3; ++r)
                {
                    st3(o + 11 * r, rw[r]->J_v);
                    st3(o + 11 * r + 3, rw[r]->J_w_a);
                    st3(o + 11 * r + 6, rw[r]->J_w_b);
                    o[11 * r + 9] = rw[r]->M_eff;
                    o[11 * r + 10] = rw[r]->bias;
                }
                o[33] = p->friction_coeff;
                o[34] = p->inv_m_11;
                o[35] = p->inv_m_12;
                o[36] = p->inv_m_22;
                o[37] = p->accumulated[0];
                o[38] = p->accumulated[1];
                o[39] = p->accumulated[2];
                keys_out[k] = key;
                point_out[k] = static_cast<uint32_t>(j);
            }
            ++k;
        }
    }
    return k;
}

// ------------------------------- dynamic_bvh handle API ---------------------------------------
void *pko_bvh_create() { return new dynamic_bvh(); }
void pko_bvh_destroy(void *t) { delete static_cast<dynamic_bvh *>(t); }
uint32_t pko_bvh_add(void *t, uint32_t id, const double *box6)
{
    return static_cast<dynamic_bvh *>(t)->add(id, aabb{ld3(box6), ld3(box6 + 3)});
}
void pko_bvh_remove(void *t, uint32_t leaf) { static_cast<dynamic_bvh *>(t)->remove_leaf(leaf); }
int pko_bvh_update(void *t, uint32_t leaf, const double *box6, const double *disp3)
{
    return static_cast<dynamic_bvh *>(t)->update_leaf(leaf, aabb{ld3(box6), ld3(box6 + 3)}, ld3(disp3)) ? 1 : 0;
}
void pko_bvh_bounds(void *t, uint32_t leaf, double *out6)
{
    const aabb &b = static_cast<dynamic_bvh *>(t)->bounds(leaf);
    st3(out6, b.min);
    st3(out6 + 3, b.max);
}
uint32_t pko_bvh_data(void *t, uint32_t leaf) { return static_cast<dynamic_bvh *>(t)->data(leaf); }
int pko_bvh_validate(void *t) { return static_cast<dynamic_bvh *>(t)->validate() ? 1 : 0; }
// Collect up to cap ids in callback order; stop_after > 0 makes the callback return false after
// that many hits (early termination, tests/dynamic_bvh/main.cpp:262-277).
uint64_t pko_bvh_query(void *t, const double *box6, uint32_t *out, uint64_t cap, uint64_t stop_after)
{
    uint64_t n = 0;
    static_cast<dynamic_bvh *>(t)->query_aabb(aabb{ld3(box6), ld3(box6 + 3)},
                                              [&](uint32_t id)
                                              {
                                                  if (n < cap) out[n] = id;
                                                  ++n;
                                                  return !(stop_after && n >= stop_after);
                                              });
    return n;
}

// ------------------------------- integrator -------------------------------------
// The per-body loops of world::step_impl (src/world.cpp:22-34 and 50-55) on flat arrays, in place.
// inertia9 / inv_inertia9: local tensors, row-major (inv_inertia9 = NULL: Matrix3d::inverse restated; a body
// with inv_mass == 0 gets a zero inverse tensor, particle.h:25-26).  flags as pk_bodies_upload.
// phase 0: loop A — apply_force(gravity·mass), integrate_vel, disp_out = vel·dt, clear_forces.
// phase 1: loop B — integrate_pos.
void pko_dynamics_step(uint64_t n, double *pos, double *quat_xyzw, double *vel, double *ang_vel, double *acc, double *torque,
                       const double *mass, const double *inertia9, const double *inv_inertia9, const uint8_t *flags,
                       const double *gravity3, double dt, int phase, double *disp_out)
{
    for (uint64_t i = 0; i < n; ++i)
    {
        if (!(flags[i] & 2) || (flags[i] & 1)) continue; // slot.available() || is_static()
        rigid_state o;
        o.pos = ld3(pos + 3 * i);
        o.q = ldq(quat_xyzw + 4 * i);
        o.vel = ld3(vel + 3 * i);
        o.ang_vel = ld3(ang_vel + 3 * i);
        o.acc = ld3(acc + 3 * i);
        o.torque = ld3(torque + 3 * i);
        o.mass = mass[i];
        o.inv_mass = 1.0 / mass[i];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) o.inertia_local.m[r][c] = inertia9[9 * i + 3 * r + c];
        if (inv_inertia9)
        {
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) o.inv_inertia_local.m[r][c] = inv_inertia9[9 * i + 3 * r + c];
        }
        else
            o.inv_inertia_local = inverse(o.inertia_local);
        if (o.inv_mass == 0.0) o.inv_inertia_local = m3{};
        o.update_derived_state();
        if (phase == 0)
        {
            v3 d = step_velocity(o, ld3(gravity3), dt);
            if (disp_out) st3(disp_out + 3 * i, d);
        }
        else
            step_position(o, dt);
        st3(pos + 3 * i, o.pos);
        quat_xyzw[4 * i + 0] = o.q.x;
        quat_xyzw[4 * i + 1] = o.q.y;
        quat_xyzw[4 * i + 2] = o.q.z;
        quat_xyzw[4 * i + 3] = o.q.w;
        st3(vel + 3 * i, o.vel);
        st3(ang_vel + 3 * i, o.ang_vel);
        st3(acc + 3 * i, o.acc);
        st3(torque + 3 * i, o.torque);
    }
}

// ------------------------------- ray casts --------------------------------------
// ray::intersect_distance (bvh.h:59-98).  dir3 is normalised by the ray constructor.  Returns 1 on a hit.
int pko_ray_box(const double *origin3, const double *dir3, const double *box6, double max_distance, double *dist_out)
{
    ray r(ld3(origin3), ld3(dir3));
    auto d = r.intersect_distance(aabb{ld3(box6), ld3(box6 + 3)}, max_distance);
    if (d && dist_out) *dist_out = *d;
    return d ? 1 : 0;
}
// dynamic_bvh::raycast.  closest = 0: generator form (bvh.h:400-450), every entered leaf in yield order.
// closest = 1: callback form (bvh.h:346-398) driven as a closest-leaf search: the callback shrinks
// max_distance to the entry distance of the leaf it was given.
// closest = 2: callback form, the callback returns 0 m: the cast ends after the first leaf (bvh.h:372).
uint64_t pko_bvh_raycast(void *t, const double *origin3, const double *dir3, double max_distance, int closest, uint32_t *ids,
                         double *dists, uint64_t cap)
{
    const dynamic_bvh &tree = *static_cast<dynamic_bvh *>(t);
    ray r(ld3(origin3), ld3(dir3));
    if (!closest)
    {
        auto v = tree.raycast(r, max_distance);
        for (uint64_t i = 0; i < v.size() && i < cap; ++i)
        {
            ids[i] = v[i].first;
            dists[i] = v[i].second;
        }
        return v.size();
    }
    uint64_t n = 0;
    tree.raycast(r, max_distance,
                 [&](uint32_t id, double d, double md)
                 {
                     if (n < cap)
                     {
                         ids[n] = id;
                         dists[n] = d;
                     }
                     ++n;
                     (void)md;
                     if (closest == 2) return 0.0;
                     return d > 0.0 ? d : std::numeric_limits<double>::min(); // 0 would end the cast (bvh.h:372)
                 });
    return n;
}

// ------------------------------- broad_phase handle API ---------------------------------------
void *pko_bp_create() { return new broad_phase(); }
void pko_bp_destroy(void *b) { delete static_cast<broad_phase *>(b); }
void pko_bp_add(void *b, uint32_t id, const double *box6, int is_static)
{
    static_cast<broad_phase *>(b)->add(id, aabb{ld3(box6), ld3(box6 + 3)}, is_static != 0);
}
void pko_bp_remove(void *b, uint32_t id) { static_cast<broad_phase *>(b)->remove(id); }
int pko_bp_update(void *b, uint32_t id, const double *box6, const double *disp3)
{
    return static_cast<broad_phase *>(b)->update_node(id, aabb{ld3(box6), ld3(box6 + 3)}, ld3(disp3)) ? 1 : 0;
}
void pko_bp_calculate(void *b) { static_cast<broad_phase *>(b)->calculate_pairs(); }
uint64_t pko_bp_pairs(void *b, uint64_t *out, uint64_t cap)
{
    auto v = static_cast<broad_phase *>(b)->sorted_pairs();
    for (uint64_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i];
    return v.size();
}
void pko_bp_stored(void *b, uint32_t id, double *out6)
{
    const aabb &bx = static_cast<broad_phase *>(b)->stored(id);
    st3(out6, bx.min);
    st3(out6 + 3, bx.max);
}

// ------------------------------- world-level collision stage ----------------------------------
// The three ★ calls of world::step_impl (src/world.cpp:30-46) driven from flat body arrays with the
// same meaning as pk_bodies_upload (include/pk_collide.h): flags bit0 = static, bit1 = alive.
// A body that becomes alive is create_rigid()'d (core/world.h:202-208: broad.add with the exact
// instance bounds, not marked moved); a body that stops being alive is remove_rigid()'d.
struct pko_world
{
    broad_phase bp;
    std::vector<uint8_t> alive;
    std::vector<uint8_t> stat;
};
void *pko_world_create() { return new pko_world(); }
void pko_world_destroy(void *w) { delete static_cast<pko_world *>(w); }

// Returns the number of leaves re-inserted this step (size of M_moved before calculate_pairs).
uint64_t pko_world_step(void *wp, const pko_shape_desc *shapes, const double *pos,
                        const double *quat_xyzw, const double *disp, const uint32_t *shape_id,
                        const uint8_t *flags, uint64_t n)
{
    pko_world &w = *static_cast<pko_world *>(wp);
    if (w.alive.size() < n)
    {
        w.alive.resize(n, 0);
        w.stat.resize(n, 0);
    }
    // flush_commands(): destroy first, then create (core/world.h:369-374; order between the two
    // does not matter for disjoint slots).
    for (uint64_t i = 0; i < w.alive.size(); ++i)
    {
        bool now = i < n && (flags[i] & 2);
        if (w.alive[i] && !now)
        {
            w.bp.remove(static_cast<uint32_t>(i));
            w.alive[i] = 0;
        }
    }
    std::vector<uint8_t> fresh(n, 0);
    for (uint64_t i = 0; i < n; ++i)
    {
        if ((flags[i] & 2) && !w.alive[i])
        {
            aabb b = body_bounds(shapes[shape_id[i]], pos + 3 * i, quat_xyzw + 4 * i);
            w.bp.add(static_cast<uint32_t>(i), b, flags[i] & 1);
            w.alive[i] = 1;
            w.stat[i] = flags[i] & 1;
            fresh[i] = 1;
        }
    }
    // step_impl loop A (src/world.cpp:22-34): every dynamic body calls update_node — including
    // one created this very step (its true box equals its stored box, so nothing moves).
    uint64_t moved = 0;
    for (uint64_t i = 0; i < n; ++i)
    {
        if (!w.alive[i] || w.stat[i]) continue;
        aabb b = body_bounds(shapes[shape_id[i]], pos + 3 * i, quat_xyzw + 4 * i);
        if (w.bp.update_node(static_cast<uint32_t>(i), b, ld3(disp + 3 * i))) ++moved;
    }
    w.bp.calculate_pairs();
    return moved;
}
uint64_t pko_world_pairs(void *wp, uint64_t *out, uint64_t cap)
{
    return pko_bp_pairs(&static_cast<pko_world *>(wp)->bp, out, cap);
}
void pko_world_stored(void *wp, uint32_t id, double *out6)
{
    pko_bp_stored(&static_cast<pko_world *>(wp)->bp, id, out6);
}

// world_base::raycast (core/world.h:260-319) over the world's two trees, merged by distance.
uint64_t pko_world_raycast(void *wp, const double *origin3, const double *dir3, double max_distance, uint32_t *ids, double *dists,
                           uint64_t cap)
{
    ray r(ld3(origin3), ld3(dir3));
    auto v = static_cast<pko_world *>(wp)->bp.raycast(r, max_distance);
    for (uint64_t i = 0; i < v.size() && i < cap; ++i)
    {
        ids[i] = v[i].first;
        dists[i] = v[i].second;
    }
    return v.size();
}

// ------------------------------- static-pose ("query") mode -----------------------------------
// The way tests/dynamic_bvh/main.cpp:600-635 drives the tree: tree.add(i, exact box) for all i,
// then query_aabb(box_i) for all i; the pair set is {(i<j) : reported}.  Returns the pair count;
// writes sorted keys.
uint64_t pko_query_pairs(const double *boxes6, uint64_t n, uint64_t *out, uint64_t cap)
{
    dynamic_bvh tree(2 * n + 16);
    std::vector<uint32_t> leaf(n);
    for (uint64_t i = 0; i < n; ++i)
        leaf[i] = tree.add(static_cast<uint32_t>(i), aabb{ld3(boxes6 + 6 * i), ld3(boxes6 + 6 * i + 3)});
    std::vector<uint64_t> keys;
    for (uint64_t i = 0; i < n; ++i)
    {
        tree.query_aabb(tree.bounds(leaf[i]),
                        [&](uint32_t j)
                        {
                            if (j > i) keys.push_back(make_pair_key(static_cast<uint32_t>(i), j));
                            return true;
                        });
    }
    std::sort(keys.begin(), keys.end());
    for (uint64_t i = 0; i < keys.size() && i < cap; ++i) out[i] = keys[i];
    return keys.size();
}

// Brute-force O(n²) reference of the same set, straight from aabb::intersects (bounds.h:87-92).
uint64_t pko_brute_pairs(const double *boxes6, uint64_t n, uint64_t *out, uint64_t cap)
{
    uint64_t cnt = 0;
    for (uint64_t i = 0; i < n; ++i)
    {
        aabb bi{ld3(boxes6 + 6 * i), ld3(boxes6 + 6 * i + 3)};
        for (uint64_t j = i + 1; j < n; ++j)
        {
            aabb bj{ld3(boxes6 + 6 * j), ld3(boxes6 + 6 * j + 3)};
            if (bi.intersects(bj))
            {
                if (cnt < cap) out[cnt] = make_pair_key(static_cast<uint32_t>(i), static_cast<uint32_t>(j));
                ++cnt;
            }
        }
    }
    return cnt;
}

int pko_max_threads()
{
    unsigned n = std::thread::hardware_concurrency();
    return n ? static_cast<int>(n) : 1;
}

} // extern "C"
