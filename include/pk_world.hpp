// pk_world.hpp — dependency-free C++17 host shim over the C ABI (pk_collide.h).
//
// Mirrors the slice of the reference's API that feeds and consumes the collision stage, with the
// same names and argument meaning, so that code written against physkit::world reads the same:
//
//   pk::world::create_rigid / remove_rigid      ↔ world_base::create_rigid / remove_rigid   core/world.h:202-218
//   pk::world::step(dt-scaled displacements)    ↔ the ★ calls of world::step_impl           src/world.cpp:30-46
//   pk::world::active_pairs()                   ↔ pair_manager::active_pairs() (sorted)      collision_phases.h:54
//   pk::world::contacts()                       ↔ one collision_info per colliding pair     collision.h:52-59
//   pk::world::enable_manifolds / manifolds()   ↔ narrow_phase::calculate + manifolds()     collision_phases.h:244-322
//   pk::world::collisions_began / _ended        ↔ the on_coll_beg / on_coll_end callbacks   collision_phases.h:314-318
//   pk::make_pair_key / extract_ids             ↔ pair_manager::make_pair_key / extract_ids collision_phases.h:56-69
//   pk::gjk_epa(a, b)                           ↔ physkit::gjk_epa                          collision.h:61-62
//
// The reference's own types (mp-units quantities over Eigen) cannot be used here because neither
// library exists in this image; INTEGRATION.md shows the gpu_world : physkit::world_base subclass
// that a PhysKit maintainer would write with them on top of the same C calls.
//
// Error behaviour follows the reference (exceptions for misuse: core/object.h:121,
// src/collision.cpp:537): every non-zero pk_status becomes a std::runtime_error carrying
// pk_strerror + pk_last_error.  There is no CPU fallback: constructing a world without a CUDA
// device throws.
#pragma once

#include "pk_collide.h"

#include <algorithm>
#include <array>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace pk
{

using vec3 = std::array<double, 3>;
using quat = std::array<double, 4>; // x, y, z, w (Eigen coeffs order, lin_alg.h:388)

struct collision_info // collision.h:52-59
{
    vec3 normal{}, world_a{}, world_b{};
    double depth{};
};

struct distance_info // no upstream counterpart: gjk_collision is boolean (src/collision.cpp:165-189); pk_gjk_distance_batch
{
    double distance{};
    vec3 closest_a{}, closest_b{};
};

struct contact_point // collision_phases.h:75-88
{
    vec3 normal{}, local_a{}, local_b{};
    double depth{};
};
struct contact_info // manifold::contact_info, collision_phases.h:93-99
{
    contact_point point;
    double normal_impulse{};
    std::array<double, 2> tangent_impulses{};
};
struct manifold_info // narrow_phase::manifold_info, collision_phases.h:203-209
{
    std::uint32_t a{}, b{};
    std::vector<contact_info> contacts; // 1..4 (manifold::max_contact_points)
};

constexpr std::uint64_t make_pair_key(std::uint32_t a, std::uint32_t b)
{
    return (static_cast<std::uint64_t>(a < b ? a : b) << 32) | (a < b ? b : a);
}
constexpr std::pair<std::uint32_t, std::uint32_t> extract_ids(std::uint64_t key)
{
    return {static_cast<std::uint32_t>(key >> 32), static_cast<std::uint32_t>(key & 0xFFFFFFFFu)};
}

class error : public std::runtime_error
{
public:
    error(int status, const std::string &what) : std::runtime_error(what), status_(status) {}
    int status() const { return status_; }

private:
    int status_;
};

struct world_desc // core/world.h:23-49 (only the capacities matter to the collision stage)
{
    std::uint32_t max_bodies = 1u << 16;
    std::uint64_t max_pairs = 1u << 20;
    std::uint64_t max_hull_vertices = 1u << 16;
    int device = 0;
    pk_mode mode = PK_MODE_WORLD;
};

class world
{
public:
    using handle = std::uint32_t; // arena slot index = body id (core/world.h:205-206)

    explicit world(const world_desc &d = {})
    {
        pk_config c{};
        c.device = d.device;
        c.mode = d.mode;
        c.max_bodies = d.max_bodies;
        c.max_shapes = d.max_bodies;
        c.max_pairs = d.max_pairs;
        c.max_hull_vertices = d.max_hull_vertices;
        int s = pk_create(&c, &ctx_);
        if (s != PK_OK) throw error(s, std::string("pk_create: ") + pk_strerror(s));
        pos_.reserve(d.max_bodies * 3);
    }
    ~world()
    {
        if (ctx_) pk_destroy(ctx_);
    }
    world(const world &) = delete;
    world &operator=(const world &) = delete;

    // shapes (what object_desc::mesh would carry, core/object.h:61-65)
    std::uint32_t shape_box(const vec3 &half) { return shape([&](std::uint32_t *id) { return pk_shape_box(ctx_, half.data(), id); }); }
    std::uint32_t shape_sphere(double r) { return shape([&](std::uint32_t *id) { return pk_shape_sphere(ctx_, r, id); }); }
    std::uint32_t shape_hull(const std::vector<vec3> &verts)
    {
        return shape([&](std::uint32_t *id)
                     { return pk_shape_hull(ctx_, verts.empty() ? nullptr : verts[0].data(), static_cast<std::uint32_t>(verts.size()), id); });
    }

    // world_base::create_rigid: the new body takes the lowest free slot, is added to the broadphase
    // with its exact bounds and is NOT marked moved (collision_phases.h:342-346).
    handle create_rigid(std::uint32_t shape_id, const vec3 &pos, const quat &orientation = {0, 0, 0, 1}, bool is_static = false)
    {
        handle h;
        if (!free_.empty())
        {
            h = free_.back();
            free_.pop_back();
        }
        else
        {
            h = static_cast<handle>(flags_.size());
            pos_.resize(pos_.size() + 3);
            quat_.resize(quat_.size() + 4);
            disp_.resize(disp_.size() + 3);
            shape_.push_back(0);
            flags_.push_back(0);
        }
        set_pose(h, pos, orientation);
        shape_[h] = shape_id;
        flags_[h] = static_cast<std::uint8_t>(2u | (is_static ? 1u : 0u));
        dirty_ = true;
        return h;
    }
    void remove_rigid(handle h)
    {
        flags_.at(h) = 0;
        free_.push_back(h);
        dirty_ = true;
    }
    void set_pose(handle h, const vec3 &pos, const quat &orientation)
    {
        for (int k = 0; k < 3; ++k) pos_[3 * h + k] = pos[k];
        for (int k = 0; k < 4; ++k) quat_[4 * h + k] = orientation[k];
    }
    // displacement = vel * dt, the predictive expansion handed to update_node (src/world.cpp:30-31)
    void set_displacement(handle h, const vec3 &d)
    {
        for (int k = 0; k < 3; ++k) disp_[3 * h + k] = d[k];
    }

    // The collision stage of one step.  Returns the counters of the step.
    pk_step_result step()
    {
        const std::uint32_t n = static_cast<std::uint32_t>(flags_.size());
        check(pk_bodies_resize(ctx_, n), "pk_bodies_resize");
        if (n)
        {
            if (dirty_)
                check(pk_bodies_upload(ctx_, pos_.data(), quat_.data(), disp_.data(), shape_.data(), flags_.data(), nullptr, 0, n),
                      "pk_bodies_upload");
            else
                check(pk_bodies_update_pose(ctx_, pos_.data(), quat_.data(), disp_.data(), 0, n), "pk_bodies_update_pose");
        }
        dirty_ = false;
        pk_step_result r{};
        check(pk_collide(ctx_, &r), "pk_collide");
        if (manifolds_) check(pk_manifolds_update(ctx_, nullptr), "pk_manifolds_update");
        return r;
    }

    // narrow_phase keeps a manifold per pair and merges every step's contact into it; with this switched on
    // step() runs that merge on the device after the collision stage
    void enable_manifolds(std::uint64_t capacity)
    {
        check(pk_manifolds_enable(ctx_, capacity), "pk_manifolds_enable");
        manifolds_ = true;
    }
    // narrow_phase::manifolds(): the non-empty ones, sorted by pair key
    std::vector<manifold_info> manifolds() const
    {
        const pk_manifold *m = nullptr;
        std::uint64_t n = 0;
        check(pk_manifolds(ctx_, &m, &n), "pk_manifolds");
        std::vector<manifold_info> out(n);
        for (std::uint64_t i = 0; i < n; ++i)
        {
            auto ids = extract_ids(m[i].key);
            out[i].a = ids.first;
            out[i].b = ids.second;
            out[i].contacts.resize(m[i].count);
            for (std::uint32_t j = 0; j < m[i].count; ++j)
            {
                const pk_manifold_point &p = m[i].points[j];
                contact_info &c = out[i].contacts[j];
                for (int k = 0; k < 3; ++k)
                {
                    c.point.normal[k] = p.normal[k];
                    c.point.local_a[k] = p.local_a[k];
                    c.point.local_b[k] = p.local_b[k];
                }
                c.point.depth = p.depth;
                c.normal_impulse = p.normal_impulse;
                c.tangent_impulses = {p.tangent_impulses[0], p.tangent_impulses[1]};
            }
        }
        return out;
    }
    // keys for which the reference would have called on_coll_beg / on_coll_end in this step
    std::vector<std::uint64_t> collisions_began() const { return events(true); }
    std::vector<std::uint64_t> collisions_ended() const { return events(false); }

    std::vector<std::uint64_t> active_pairs() const
    {
        const std::uint64_t *k = nullptr;
        std::uint64_t n = 0;
        check(pk_pairs(ctx_, &k, &n), "pk_pairs");
        return std::vector<std::uint64_t>(k, k + n);
    }
    std::vector<std::pair<std::uint64_t, collision_info>> contacts() const
    {
        const pk_contact *c = nullptr;
        std::uint64_t n = 0;
        check(pk_contacts(ctx_, &c, &n), "pk_contacts");
        std::vector<std::pair<std::uint64_t, collision_info>> out(n);
        for (std::uint64_t i = 0; i < n; ++i)
        {
            out[i].first = c[i].key;
            for (int k = 0; k < 3; ++k)
            {
                out[i].second.normal[k] = c[i].normal[k];
                out[i].second.world_a[k] = c[i].world_a[k];
                out[i].second.world_b[k] = c[i].world_b[k];
            }
            out[i].second.depth = c[i].depth;
        }
        return out;
    }

    // world_base::raycast(ray, max_dist) (core/world.h:260-319): every body whose stored box the ray enters within
    // max_dist, with the entry distance, ordered by (distance, handle).  (The reference yields each tree's
    // entries in traversal order and merges the two streams by distance.)
    std::vector<std::pair<handle, double>> raycast(const vec3 &origin, const vec3 &direction, double max_dist) const
    {
        std::vector<pk_ray_hit> buf(64);
        std::uint64_t n = 0;
        int s = pk_raycast(ctx_, origin.data(), direction.data(), &max_dist, nullptr, 1, PK_RAY_ALL, buf.data(), buf.size(), &n);
        if (s == PK_E_PAIR_OVERFLOW)
        {
            buf.resize(n);
            s = pk_raycast(ctx_, origin.data(), direction.data(), &max_dist, nullptr, 1, PK_RAY_ALL, buf.data(), buf.size(), &n);
        }
        check(s, "pk_raycast");
        std::vector<std::pair<handle, double>> out(n);
        for (std::uint64_t i = 0; i < n; ++i) out[i] = {buf[i].body, buf[i].distance};
        std::sort(out.begin(), out.end(), [](const auto &a, const auto &b) { return a.second < b.second || (a.second == b.second && a.first < b.first); });
        return out;
    }

    // physkit::gjk_epa for two bodies of this world (argument order as given, not min/max)
    std::optional<collision_info> gjk_epa(handle a, handle b)
    {
        pk_contact c{};
        std::uint8_t hit = 0;
        check(pk_bodies_resize(ctx_, static_cast<std::uint32_t>(flags_.size())), "pk_bodies_resize");
        check(pk_bodies_upload(ctx_, pos_.data(), quat_.data(), disp_.data(), shape_.data(), flags_.data(), nullptr, 0,
                               static_cast<std::uint32_t>(flags_.size())),
              "pk_bodies_upload");
        check(pk_gjk_epa_batch(ctx_, &a, &b, 1, &c, &hit), "pk_gjk_epa_batch");
        if (!hit) return std::nullopt;
        collision_info r;
        for (int k = 0; k < 3; ++k) r.normal[k] = c.normal[k], r.world_a[k] = c.world_a[k], r.world_b[k] = c.world_b[k];
        r.depth = c.depth;
        return r;
    }

    // closest distance and closest points of two bodies of this world; nullopt when they touch or overlap (then gjk_epa
    // has the depth).  The reference has no such query; BASELINE.json's north_star names it.
    std::optional<distance_info> distance(handle a, handle b)
    {
        pk_distance d{};
        std::uint8_t sep = 0;
        check(pk_bodies_resize(ctx_, static_cast<std::uint32_t>(flags_.size())), "pk_bodies_resize");
        check(pk_bodies_upload(ctx_, pos_.data(), quat_.data(), disp_.data(), shape_.data(), flags_.data(), nullptr, 0,
                               static_cast<std::uint32_t>(flags_.size())),
              "pk_bodies_upload");
        check(pk_gjk_distance_batch(ctx_, &a, &b, 1, &d, &sep), "pk_gjk_distance_batch");
        if (!sep) return std::nullopt;
        distance_info r;
        r.distance = d.distance;
        for (int k = 0; k < 3; ++k) r.closest_a[k] = d.point_a[k], r.closest_b[k] = d.point_b[k];
        return r;
    }

    pk_ctx *native() const { return ctx_; }

private:
    std::vector<std::uint64_t> events(bool began) const
    {
        const std::uint64_t *b = nullptr, *e = nullptr;
        std::uint64_t nb = 0, ne = 0;
        check(pk_manifold_events(ctx_, &b, &nb, &e, &ne), "pk_manifold_events");
        return began ? std::vector<std::uint64_t>(b, b + nb) : std::vector<std::uint64_t>(e, e + ne);
    }
    bool manifolds_ = false;

private:
    template <typename F> std::uint32_t shape(F &&f)
    {
        std::uint32_t id = 0;
        check(f(&id), "pk_shape_*");
        return id;
    }
    void check(int s, const char *what) const
    {
        if (s != PK_OK) throw error(s, std::string(what) + ": " + pk_strerror(s) + " (" + pk_last_error(ctx_) + ")");
    }

    pk_ctx *ctx_ = nullptr;
    std::vector<double> pos_, quat_, disp_;
    std::vector<std::uint32_t> shape_;
    std::vector<std::uint8_t> flags_;
    std::vector<handle> free_;
    bool dirty_ = true;
};

} // namespace pk
