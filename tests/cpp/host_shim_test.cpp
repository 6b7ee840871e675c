// Exercises include/pk_world.hpp the way the reference's tests/co/co_tests.cpp:358-382 drives
// physkit::world: two boxes approach each other; the pair must appear one step after the first move
// (first-step quirk) and the contact must have a unit normal and positive depth once they overlap.
// Exit code: 0 ok, 3 no CUDA device (expected on the CPU-only build box), 1 failure.
#include "pk_world.hpp"

#include <cmath>
#include <cstdio>

int main()
{
    try
    {
        pk::world_desc d;
        d.max_bodies = 64;
        d.max_pairs = 1024;
        pk::world w(d);
        w.enable_manifolds(256);
        auto box = w.shape_box({0.5, 0.5, 0.5});
        auto a = w.create_rigid(box, {-2.0, 0.0, 0.0});
        auto b = w.create_rigid(box, {2.0, 0.0, 0.0});
        auto ground = w.create_rigid(w.shape_box({50.0, 0.5, 50.0}), {0.0, -5.0, 0.0}, {0, 0, 0, 1}, true);
        const double dt = 1.0 / 60.0, v = 3.0;
        bool seen_pair = false, seen_contact = false, seen_begin = false;
        std::size_t max_points = 0;
        double xa = -2.0, xb = 2.0;
        for (int step = 0; step < 36; ++step) // stop at 0.5 m overlap: the x axis is still the unique shallowest one
        {
            w.set_pose(a, {xa, 0.0, 0.0}, {0, 0, 0, 1});
            w.set_pose(b, {xb, 0.0, 0.0}, {0, 0, 0, 1});
            w.set_displacement(a, {v * dt, 0.0, 0.0});
            w.set_displacement(b, {-v * dt, 0.0, 0.0});
            auto r = w.step();
            if (step == 0 && r.num_pairs != 0) return std::printf("FAIL: pairs on the first step\n"), 1;
            auto pairs = w.active_pairs();
            for (auto k : pairs)
                if (k == pk::make_pair_key(a, b)) seen_pair = true;
            for (auto &c : w.contacts())
            {
                if (c.first != pk::make_pair_key(a, b)) continue;
                double n = std::sqrt(c.second.normal[0] * c.second.normal[0] + c.second.normal[1] * c.second.normal[1] +
                                     c.second.normal[2] * c.second.normal[2]);
                if (std::fabs(n - 1.0) > 1e-6 || !(c.second.depth > 0.0)) return std::printf("FAIL: bad contact\n"), 1;
                if (std::fabs(std::fabs(c.second.normal[0]) - 1.0) > 1e-9) return std::printf("FAIL: normal not along x\n"), 1;
                seen_contact = true;
            }
            // narrow_phase::calculate: the pair's manifold starts with on_coll_beg and gathers points as the
            // boxes slide into each other (collision_phases.h:244-320)
            for (auto k : w.collisions_began())
                if (k == pk::make_pair_key(a, b)) seen_begin = true;
            for (auto &m : w.manifolds())
                if (pk::make_pair_key(m.a, m.b) == pk::make_pair_key(a, b))
                {
                    if (m.contacts.empty() || m.contacts.size() > 4) return std::printf("FAIL: manifold size\n"), 1;
                    if (m.contacts.size() > max_points) max_points = m.contacts.size();
                }
            xa += v * dt;
            xb -= v * dt;
        }
        if (!seen_begin || max_points < 1) return std::printf("FAIL: manifold begin %d points %zu\n", seen_begin, max_points), 1;
        // world_base::raycast over the last step's tree: a ray along +x from far left enters a's fat box first, then b's
        auto rc = w.raycast({-30.0, 0.0, 0.0}, {1.0, 0.0, 0.0}, 100.0);
        if (rc.size() != 2 || rc[0].first != a || rc[1].first != b || !(rc[0].second > 20.0 && rc[0].second <= rc[1].second))
            return std::printf("FAIL: raycast %zu\n", rc.size()), 1;
        if (!w.raycast({-30.0, 20.0, 0.0}, {1.0, 0.0, 0.0}, 100.0).empty()) return std::printf("FAIL: raycast miss\n"), 1;
        auto one = w.gjk_epa(a, b);
        if (!seen_pair || !seen_contact || !one) return std::printf("FAIL: pair %d contact %d gjk %d\n", seen_pair, seen_contact, (int)one.has_value()), 1;
        // distance query (not in the reference): box a (half 0.5, y = 0) hangs 4 m above the ground slab's top (y = −4.5);
        // the two overlapping boxes have no distance
        auto da = w.distance(a, ground);
        if (!da || std::fabs(da->distance - 4.0) > 1e-12 || std::fabs(da->closest_a[1] + 0.5) > 1e-12 || std::fabs(da->closest_b[1] + 4.5) > 1e-12)
            return std::printf("FAIL: distance %f\n", da ? da->distance : -1.0), 1;
        if (w.distance(a, b)) return std::printf("FAIL: distance of overlapping boxes\n"), 1;
        std::printf("host shim ok\n");
        return 0;
    }
    catch (const pk::error &e)
    {
        if (e.status() == PK_E_NO_DEVICE) return std::printf("no CUDA device: %s\n", e.what()), 3;
        std::printf("FAIL: %s\n", e.what());
        return 1;
    }
}
