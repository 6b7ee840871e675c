// pk_manifold.cuh — narrow_phase's contact manifolds on the device (reference
// include/physkit/collision/collision_phases.h:75-327; SURVEY §8f-1, the first consumer of gjk_epa's result).
//
// The reference keeps one manifold (≤4 contact points with cached solver impulses) per pair of the pair set
// and, every step, merges the pair's new contact into it: warm-start match, drift / breaking test of the old
// points under the bodies' new poses, add_reduce when a fifth point arrives (narrow_phase::calculate,
// :244-320).  Only non-empty manifolds carry state.  Here they live in an array sorted by pair key, next to
// the step's sorted pair keys and contacts, so no hash map is needed:
//   manifold_old_kernel   one thread per manifold of the previous step: binary search of its key in the
//                         step's pair list (gone → dropped, on_pair_removed :225-242), merge with the pair's
//                         contact if it has one, stage the result, mark the contact consumed
//   manifold_new_kernel   one thread per contact: contacts nobody consumed start a manifold (on_coll_beg)
//   the staged (key, index) candidates are radix-sorted by key and gathered into the next sorted array.
// Arithmetic follows the reference statement by statement (Eigen operation order of pk_common.cuh, no FMA), so
// results are bit-identical to the oracle's restatement (oracle/pk_oracle.hpp: manifold_merge).
#pragma once

#include "pk_common.cuh"

namespace pk
{

// manifold::contact_info (:93-99) with contact_point (:75-88) flattened: 13 doubles
struct ManifoldPoint
{
    double normal[3];
    double local_a[3];
    double local_b[3];
    double depth;
    double normal_impulse;
    double tangent_impulses[2];
};
static_assert(sizeof(ManifoldPoint) == 104, "ManifoldPoint mirrors pk_manifold_point");

struct ManifoldRec // mirrors pk_manifold
{
    uint64_t key;
    uint32_t count;
    uint32_t _pad;
    ManifoldPoint pt[4];
};
static_assert(sizeof(ManifoldRec) == 432, "ManifoldRec mirrors pk_manifold");

__device__ __forceinline__ d3 mp_ld(const double *p) { return {p[0], p[1], p[2]}; }
__device__ __forceinline__ void mp_st(double *p, d3 v)
{
    p[0] = v.x;
    p[1] = v.y;
    p[2] = v.z;
}

// manifold::add_reduce (:139-198): of five points keep the deepest, the one farthest from it, the one spanning
// the largest triangle with those two, and the one farthest from the third.
__device__ __forceinline__ void manifold_add_reduce(ManifoldRec &m, const ManifoldPoint &np)
{
    ManifoldPoint pool[5];
    for (int i = 0; i < 4; ++i) pool[i] = m.pt[i];
    pool[4] = np;
    int best[4] = {0, 1, 2, 3};
    for (int i = 1; i < 5; ++i)
        if (pool[i].depth > pool[best[0]].depth) best[0] = i;
    double max_dist2 = -1.0;
    for (int i = 0; i < 5; ++i)
    {
        if (i == best[0]) continue;
        const double dist2 = sqnorm(mp_ld(pool[i].local_a) - mp_ld(pool[best[0]].local_a));
        if (dist2 > max_dist2)
        {
            max_dist2 = dist2;
            best[1] = i;
        }
    }
    double max_area2 = -1.0;
    const d3 edge0 = mp_ld(pool[best[1]].local_a) - mp_ld(pool[best[0]].local_a);
    for (int i = 0; i < 5; ++i)
    {
        if (i == best[0] || i == best[1]) continue;
        const d3 edge1 = mp_ld(pool[i].local_a) - mp_ld(pool[best[0]].local_a);
        const double area2 = sqnorm(cross(edge0, edge1));
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
        const double dist2 = sqnorm(mp_ld(pool[i].local_a) - mp_ld(pool[best[2]].local_a));
        if (dist2 > max_dist2)
        {
            max_dist2 = dist2;
            best[3] = i;
        }
    }
    for (int k = 0; k < 4; ++k) m.pt[k] = pool[best[k]];
    m.count = 4;
}
// manifold::add_contact (:127-133)
__device__ __forceinline__ void manifold_add(ManifoldRec &m, const ManifoldPoint &p)
{
    if (m.count < 4)
        m.pt[m.count++] = p;
    else
        manifold_add_reduce(m, p);
}

// contact_point(info, a, b) (:78-82) from a contact record and the two poses
__device__ __forceinline__ ManifoldPoint manifold_point_from_contact(const ContactRec &c, d3 pos_a, dq q_a, d3 pos_b, dq q_b)
{
    ManifoldPoint p;
    p.normal[0] = c.normal[0];
    p.normal[1] = c.normal[1];
    p.normal[2] = c.normal[2];
    mp_st(p.local_a, rotate(conjugate(q_a), mp_ld(c.world_a) - pos_a)); // particle::project_to_local
    mp_st(p.local_b, rotate(conjugate(q_b), mp_ld(c.world_b) - pos_b));
    p.depth = c.depth;
    p.normal_impulse = 0.0;
    p.tangent_impulses[0] = p.tangent_impulses[1] = 0.0;
    return p;
}

// One pair's share of narrow_phase::calculate (:252-312).
__device__ __forceinline__ void manifold_merge(const ManifoldRec &old_man, bool have_new, ManifoldPoint nc, d3 pos_a, dq q_a, d3 pos_b,
                                               dq q_b, ManifoldRec &out)
{
    constexpr double distance2_eps = .005 * .005;        // :212
    constexpr double contact_breaking_threshold = 0.05; // :213
    constexpr double drift2_eps = 0.06;                 // :267
    out.key = old_man.key;
    out.count = 0;
    out._pad = 0;
    for (uint32_t k = 0; k < old_man.count; ++k)
    {
        ManifoldPoint oc = old_man.pt[k];
        if (have_new)
        {
            if (sqnorm(mp_ld(nc.local_a) - mp_ld(oc.local_a)) < distance2_eps || sqnorm(mp_ld(nc.local_b) - mp_ld(oc.local_b)) < distance2_eps)
            {
                // warm start: the new point inherits the cached impulses of an old one at the same place
                nc.normal_impulse = oc.normal_impulse;
                nc.tangent_impulses[0] = oc.tangent_impulses[0];
                nc.tangent_impulses[1] = oc.tangent_impulses[1];
                continue;
            }
        }
        const d3 world_old_a = rotate(q_a, mp_ld(oc.local_a)) + pos_a; // particle::project_to_world
        const d3 world_old_b = rotate(q_b, mp_ld(oc.local_b)) + pos_b;
        const d3 normal = have_new ? mp_ld(nc.normal) : mp_ld(oc.normal);
        const d3 relative = world_old_b - world_old_a;
        const double depth = dot(relative, normal);
        const d3 projected_a = world_old_a + normal * depth;
        const double drift2 = sqnorm(projected_a - world_old_b);
        if (depth > -contact_breaking_threshold && drift2 < drift2_eps)
        {
            oc.depth = depth;
            mp_st(oc.normal, normal);
            manifold_add(out, oc);
        }
    }
    if (have_new) manifold_add(out, nc);
}

__device__ __forceinline__ void manifold_load_pose(const double *__restrict__ pos, const double *__restrict__ quat, uint32_t body, d3 &p, dq &q)
{
    p = {pos[3ull * body], pos[3ull * body + 1], pos[3ull * body + 2]};
    q = {quat[4ull * body], quat[4ull * body + 1], quat[4ull * body + 2], quat[4ull * body + 3]};
}

// counters: [0] candidates (non-empty manifolds of the next array), [1] new manifolds staged, [2] began, [3] ended
__global__ void __launch_bounds__(128)
manifold_old_kernel(const ManifoldRec *__restrict__ prev, uint64_t m_prev, const uint64_t *__restrict__ pair_keys, uint64_t npairs,
                    const uint8_t *__restrict__ hit, const uint32_t *__restrict__ out_index, const uint8_t *__restrict__ valid,
                    const ContactRec *__restrict__ contacts, const double *__restrict__ pos, const double *__restrict__ quat,
                    ManifoldRec *__restrict__ stage, uint8_t *__restrict__ consumed, uint64_t *__restrict__ cand_keys,
                    uint32_t *__restrict__ cand_src, uint64_t cand_cap, uint64_t *__restrict__ ended,
                    unsigned long long *__restrict__ counters)
{
    const uint64_t j = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (j >= m_prev) return;
    const ManifoldRec old_man = prev[j];
    // lower bound of the key in the sorted pair list
    uint64_t lo = 0, hi = npairs;
    while (lo < hi)
    {
        const uint64_t mid = (lo + hi) >> 1;
        if (pair_keys[mid] < old_man.key)
            lo = mid + 1;
        else
            hi = mid;
    }
    if (lo >= npairs || pair_keys[lo] != old_man.key) return; // on_pair_removed: dropped without a callback
    const uint32_t ia = static_cast<uint32_t>(old_man.key >> 32), ib = static_cast<uint32_t>(old_man.key & 0xFFFFFFFFu);
    d3 pa, pb;
    dq qa, qb;
    manifold_load_pose(pos, quat, ia, pa, qa);
    manifold_load_pose(pos, quat, ib, pb, qb);
    bool have_new = false;
    ManifoldPoint nc{};
    if (hit[lo])
    {
        const uint32_t slot = out_index[lo];
        if (valid[slot])
        {
            have_new = true;
            nc = manifold_point_from_contact(contacts[slot], pa, qa, pb, qb);
            consumed[slot] = 1;
        }
    }
    ManifoldRec nm;
    manifold_merge(old_man, have_new, nc, pa, qa, pb, qb, nm);
    if (nm.count)
    {
        stage[j] = nm;
        const unsigned long long c = atomicAdd(counters + 0, 1ull);
        if (c < cand_cap) // beyond: the host reports the capacity problem from the counter
        {
            cand_keys[c] = nm.key;
            cand_src[c] = static_cast<uint32_t>(j);
        }
    }
    else
        ended[atomicAdd(counters + 3, 1ull)] = old_man.key; // on_coll_end
}

__global__ void __launch_bounds__(128)
manifold_new_kernel(const ContactRec *__restrict__ contacts, const uint8_t *__restrict__ valid, const uint8_t *__restrict__ consumed,
                    uint64_t nslots, const double *__restrict__ pos, const double *__restrict__ quat, ManifoldRec *__restrict__ stage,
                    uint64_t m_prev, uint64_t stage_cap, uint64_t *__restrict__ cand_keys, uint32_t *__restrict__ cand_src,
                    uint64_t cand_cap, uint64_t *__restrict__ began, unsigned long long *__restrict__ counters)
{
    const uint64_t s = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (s >= nslots || !valid[s] || consumed[s]) return;
    const ContactRec c = contacts[s];
    const unsigned long long idx = m_prev + atomicAdd(counters + 1, 1ull);
    if (idx >= stage_cap) return; // capacity exceeded: reported by the host from counters[1]
    const uint32_t ia = static_cast<uint32_t>(c.key >> 32), ib = static_cast<uint32_t>(c.key & 0xFFFFFFFFu);
    d3 pa, pb;
    dq qa, qb;
    manifold_load_pose(pos, quat, ia, pa, qa);
    manifold_load_pose(pos, quat, ib, pb, qb);
    ManifoldRec nm;
    nm.key = c.key;
    nm.count = 1;
    nm._pad = 0;
    nm.pt[0] = manifold_point_from_contact(c, pa, qa, pb, qb);
    for (int k = 1; k < 4; ++k) nm.pt[k] = ManifoldPoint{};
    stage[idx] = nm;
    const unsigned long long cpos = atomicAdd(counters + 0, 1ull);
    if (cpos < cand_cap)
    {
        cand_keys[cpos] = c.key;
        cand_src[cpos] = static_cast<uint32_t>(idx);
    }
    const unsigned long long bpos = atomicAdd(counters + 2, 1ull); // on_coll_beg
    if (bpos < cand_cap) began[bpos] = c.key;
}

__global__ void __launch_bounds__(128)
manifold_gather_kernel(const uint32_t *__restrict__ src, uint64_t m, const ManifoldRec *__restrict__ stage, ManifoldRec *__restrict__ next)
{
    const uint64_t r = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (r >= m) return;
    ManifoldRec rec = stage[src[r]];
    // unused points are zero in the public record
    for (uint32_t k = rec.count; k < 4; ++k) rec.pt[k] = ManifoldPoint{};
    next[r] = rec;
}

// what the constraint solver leaves behind: imp[m][4][3] = normal, tangent 0, tangent 1
__global__ void __launch_bounds__(256)
manifold_impulses_kernel(ManifoldRec *__restrict__ recs, uint64_t m, const double *__restrict__ imp)
{
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= 4 * m) return;
    ManifoldRec &r = recs[i >> 2];
    const uint32_t k = static_cast<uint32_t>(i & 3u);
    if (k >= r.count) return;
    r.pt[k].normal_impulse = imp[3 * i];
    r.pt[k].tangent_impulses[0] = imp[3 * i + 1];
    r.pt[k].tangent_impulses[1] = imp[3 * i + 2];
}

} // namespace pk
