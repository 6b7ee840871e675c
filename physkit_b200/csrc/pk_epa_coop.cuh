// pk_epa_coop.cuh — EPA (reference src/collision.cpp:251-509): one lane per pair for the sequential part of an
// iteration, the whole warp for the part that is parallel over horizon edges.
//
// ncu on epa_scan_kernel (profiles/r1b_epa_scan_fullsize.md): 13.5 of 32 lanes active, 4 400 warp instructions per
// warp iteration for 2 400 per lane.  Two things cost the lanes: idle lanes waiting for a refill, and the loops
// whose trip count differs from lane to lane — above all the face loop (one new face per horizon edge, ≈40 % of
// the instructions), which a warp runs max(h) times for a mean of h.
//
// An EPA iteration splits into
//   S1 (per pair, sequential): pop the closest face, support point, convergence test, flood fill → horizon
//      (collision.cpp:397-408, 461-470, 315-353), slot assignment for the new faces;
//   S2 (per horizon edge, independent): face plane of (start, end, p), key, link to the face across the horizon,
//      ring links to the two neighbouring new faces (collision.cpp:273-297, 305-313, 475-497).
// S1 runs one lane per pair as before.  S2 is dealt out evenly: the edges of all pairs of the warp are numbered
// by a prefix sum and lane l takes edges l, l+32, …, whoever owns them — lanes whose pair has finished, whose
// horizon is short or which have no pair at all (the tail of the launch) work for the others, and no lane
// executes anything redundantly.  The owner publishes its horizon, the new vertex and the slots in shared
// memory; the polytope stays in the owner's slab, which any lane of the warp can address.
//
// Everything else follows epa_scan_kernel: float keys + scan instead of a heap for pairs with a sphere (SCAN),
// the reference's heap restated for polyhedron pairs (HEAP), recycled face slots, hand-back of the cases that
// need the heap's history.  Finished pairs are parked and their results written when the warp refills, so that
// the result path runs for several lanes at once.  Results are bit-identical to epa_kernel.
#pragma once

#include "pk_epa_scan.cuh"

namespace pk
{

#ifndef PK_EC_MIN_BLOCKS
#define PK_EC_MIN_BLOCKS 5
#endif
#ifndef PK_EC_FETCH_MIN
#define PK_EC_FETCH_MIN 8
#endif

struct EcSmem
{
    EsPop pop;
    double f[2][10][ES_THREADS];      // shape views: p xyz, h xyz, q xyzw
    uint32_t vert_off[2][ES_THREADS]; // HULL: first vertex in the context's vertex pool
    float hull_r[2][ES_THREADS];
    int kind[2][ES_THREADS];
    uint32_t nverts[2][ES_THREADS];
    uint32_t hz[ES_HORIZON][ES_THREADS]; // start:7 | end:7 | adjacent slot:8 | new slot:8 (bits 24-31)
    double pnew[3][ES_THREADS];          // the support point of this iteration (apex of the new faces)
    uint32_t meta[ES_THREADS];           // index of that vertex:8 | creation serial of the first new face << 8
    uint32_t bad[ES_THREADS];            // set by any lane that finds the owner's horizon improper
};
static_assert(sizeof(EcSmem) <= 48 * 1024, "EcSmem must fit static shared memory");

// One new face (collision.cpp:475-482 init_face + link_faces, 484-497 ring links) of the pair owned by lane L of
// this warp; executed by whichever lane the edge was dealt to.
template <bool HEAP>
__device__ __forceinline__ void ec_make_face(EcSmem &shm, const EsSlab &osl, int ot, int e, int nh)
{
    const uint32_t hc = shm.hz[e][ot];
    const int st = hz_start(hc), en = hz_end(hc), adj = hz_adj(hc), slot = hz_slot(hc);
    const uint32_t meta = shm.meta[ot];
    const int p_idx = static_cast<int>(meta & 0xFFu);
    const d3 p{shm.pnew[0][ot], shm.pnew[1][ot], shm.pnew[2][ot]};
    const d3 cs = osl.vp(st), ce = osl.vp(en);
    const unsigned long long at = osl.topo[adj];
    bool flip_unused;
    const double4 pl = es_face_plane(cs, ce, p, false, d3{0, 0, 0}, flip_unused);
    osl.store_plane(slot, d3{pl.x, pl.y, pl.z}, pl.w);
    bool bad = !(fabs(pl.w) < 1e30); // NaN / inf: the key order would not be the heap's
    if constexpr (!HEAP)
    {
        if (slot < ES_KEYS)
            shm.pop.th[ot].key[slot] = __double2float_rd(pl.w);
        else
            osl.gkey[slot - ES_KEYS] = __double2float_rd(pl.w);
    }
    // link_faces(f, adj_face, start, end): on the old face the shared edge starts at `end` (collision.cpp:305-313)
    const int e2 = (es_v(at, 0) == en) ? 0 : (es_v(at, 1) == en ? 1 : 2);
    if (es_v(at, e2) != en) bad = true; // unmatched link: slot recycling is no longer safe
    osl.set_adj(adj, e2, slot);
    // ring: the edge that starts at `end` gives the neighbour on edge 1, the edge that ends at `start` the one on
    // edge 2.  A proper horizon has exactly one of each for every edge, no self loop and no 2-cycle (which the
    // reference links one way only); anything else is handed back.
    int succ = 0xFF, pred = 0xFF, nsucc = 0, npred = 0;
    for (int j = 0; j < nh; ++j)
    {
        const uint32_t hj = shm.hz[j][ot];
        if (hz_start(hj) == en)
        {
            succ = hz_slot(hj);
            ++nsucc;
            if (hz_end(hj) == st) bad = true;
        }
        if (hz_end(hj) == st)
        {
            pred = hz_slot(hj);
            ++npred;
        }
    }
    if (nsucc != 1 || npred != 1 || st == en) bad = true;
    osl.topo[slot] = static_cast<unsigned long long>(st) | (static_cast<unsigned long long>(en) << 8) |
                     (static_cast<unsigned long long>(p_idx) << 16) | (static_cast<unsigned long long>(adj) << 24) |
                     (static_cast<unsigned long long>(succ) << 32) | (static_cast<unsigned long long>(pred) << 40) |
                     (static_cast<unsigned long long>((meta >> 8) + static_cast<uint32_t>(e)) << 48); // creation serial
    if (bad) shm.bad[ot] = 1u;
}

// Work lists, classes, hand-back protocol and parameters as epa_scan_kernel (pk_epa_scan.cuh).
template <bool HEAP, bool MIRROR>
__global__ void __launch_bounds__(ES_THREADS, PK_EC_MIN_BLOCKS)
epa_coop_kernel(BodyArrays bodies, const uint64_t *__restrict__ keys, const uint32_t *__restrict__ pair_a,
                const uint32_t *__restrict__ pair_b, const SimplexRec *__restrict__ simplices,
                const unsigned long long *__restrict__ hit_count_ptr, uint64_t hit_capacity,
                const uint32_t *__restrict__ out_index, const uint32_t *__restrict__ order, ContactRec *__restrict__ contacts,
                uint8_t *__restrict__ valid, unsigned char *__restrict__ slabs, unsigned long long *__restrict__ cursor,
                unsigned long long *__restrict__ counters /* [0]=valid contacts, [1]=dropped: no room for the contact */,
                uint32_t *__restrict__ fallback_list, unsigned long long *__restrict__ fallback_count,
                const unsigned long long *__restrict__ class_count, const uint32_t *__restrict__ leftovers,
                const unsigned long long *__restrict__ leftover_count, const EpaInit *__restrict__ init, ContactRec *contacts_host)
{
    __shared__ EcSmem shm;
    const int t = threadIdx.x;
    const int lane = t & 31, wbase = t & ~31;
    const uint64_t tid = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    const size_t slab_bytes = es_slab_bytes(HEAP);
    unsigned char *const wslab = slabs + (tid - static_cast<uint64_t>(lane)) * slab_bytes; // slab of lane 0 of this warp
    const EsSlab sl(wslab + static_cast<size_t>(lane) * slab_bytes, HEAP);
    const EsHeapRef hp{&shm.pop.th[t], sl.hd, sl.hf};
    unsigned long long nhits = *hit_count_ptr;
    if (nhits > hit_capacity) nhits = hit_capacity;
    const unsigned long long fb_capacity = nhits;
    unsigned long long first_hit = class_count[2] + class_count[1]; // where the polyhedron pairs start
    if (first_hit > nhits) first_hit = nhits;
    unsigned long long nleft = 0;
    if (HEAP)
    {
        nhits -= first_hit;
        nleft = *leftover_count < fb_capacity ? *leftover_count : fb_capacity;
    }
    else
    {
        nhits = first_hit;
        first_hit = 0;
    }
    const float INF = es_inf();

    bool active = false, done = false;
    bool pending = false; // the pair has converged on face pend_face; its record is written at the next refill
    int pend_face = 0;
    int nverts = 0, iter = 0, hi = 0; // hi: slots [0, hi) have been used by the current polytope
    bool keys_dirty = true;           // the key area does not hold +inf beyond the current polytope
    int gdirty = ES_SLOTS;            // slab keys [ES_KEYS, gdirty) may hold something else than +inf
    unsigned long long fm0 = 0, fm1 = 0, fm2 = 0; // free slots, 64 per word
    uint32_t out_slot = 0, cur_sidx = 0;
    uint64_t key = 0;
    double stale_lb = 1e300; // see epa_scan_kernel: lower bound of the lazily deleted heap entries' distances
    int batch_n = 0;         // hz[0, batch_n) still describes the last batch of faces, in push order
    unsigned long long n_valid = 0, n_dropped = 0;
    int heap_size = 0, nfaces = 0; // nfaces: faces created so far = serial of the next one
    constexpr unsigned FULL = 0xFFFFFFFFu;
    int fb = 0; // reason + 1 when the current pair has to be handed back
    bool flush_pending = false;
    uint32_t flush_slot = 0;

    auto is_free = [&](int f) -> bool
    {
        const unsigned long long w = (f < 64) ? fm0 : (f < 128 ? fm1 : fm2);
        return (w >> (f & 63)) & 1ull;
    };
    auto kill_slot = [&](int f)
    {
        if constexpr (!HEAP)
        {
            if (f < ES_KEYS)
                shm.pop.th[t].key[f] = INF;
            else
                sl.gkey[f - ES_KEYS] = INF;
        }
        if (f < 64)
            fm0 |= 1ull << f;
        else if (f < 128)
            fm1 |= 1ull << (f - 64);
        else
            fm2 |= 1ull << (f - 128);
    };
    auto key_of = [&](int s) -> float { return (s < ES_KEYS) ? shm.pop.th[t].key[s] : sl.gkey[s - ES_KEYS]; };
    auto take_slot = [&]() -> int
    {
        int slot;
        if (HEAP && nfaces < ES_SLOTS)
            slot = nfaces;
        else if (fm0)
            slot = __ffsll(static_cast<long long>(fm0)) - 1;
        else if (fm1)
            slot = 64 + __ffsll(static_cast<long long>(fm1)) - 1;
        else
            slot = 128 + __ffsll(static_cast<long long>(fm2)) - 1;
        if (slot < 64)
            fm0 &= ~(1ull << slot);
        else if (slot < 128)
            fm1 &= ~(1ull << (slot - 64));
        else
            fm2 &= ~(1ull << (slot - 128));
        return slot;
    };

    for (;;)
    {
        if (fb)
        {
            unsigned long long i = atomicAdd(fallback_count, 1ull);
            if (i < fb_capacity) fallback_list[i] = cur_sidx; // one word: publication needs no fence
            fb = 0;
            active = false;
        }
        const unsigned m_active = __ballot_sync(FULL, active);
        const unsigned m_idle = __ballot_sync(FULL, !active && !done);
        if (m_active == 0 && m_idle == 0) break;
        const bool refill = !active && !done && (__popc(m_idle) >= PK_EC_FETCH_MIN || m_active == 0);
        if (refill && pending)
        {
            // collision.cpp:424-454 for the face the pair converged on (or the best guess after 64 iterations)
            es_write_result(sl, sl.load_plane(pend_face), sl.topo[pend_face], contacts + out_slot, key,
                            MIRROR ? reinterpret_cast<ContactRec *>(&shm.pop.th[t]) : nullptr);
            valid[out_slot] = 1;
            ++n_valid;
            pending = false;
            if constexpr (MIRROR)
            {
                flush_pending = true;
                flush_slot = out_slot;
                if (hi < 22) hi = 22; // the staged record covers the first 22 key slots: the fetch below resets them
            }
        }
        if constexpr (MIRROR)
        {
            // pk_collide: finished records go to the caller's pinned buffer from here (see epa_scan_kernel): the
            // record waits in the finishing lane's key area, eleven lanes store it with one 8-byte store each
            unsigned m_flush = __ballot_sync(FULL, flush_pending);
            if (m_flush)
            {
                __syncwarp();
                while (m_flush)
                {
                    const int L = __ffs(static_cast<int>(m_flush)) - 1;
                    m_flush &= m_flush - 1u;
                    const uint32_t slot_l = __shfl_sync(FULL, flush_slot, L);
                    if (lane < 11)
                        reinterpret_cast<unsigned long long *>(contacts_host + slot_l)[lane] =
                            reinterpret_cast<const unsigned long long *>(&shm.pop.th[wbase + L])[lane];
                }
                __syncwarp();
                flush_pending = false;
            }
        }
        if (refill)
        {
            unsigned long long slot = atomicAdd(cursor, 1ull);
            if (slot >= nleft + nhits)
                done = true;
            else
            {
                cur_sidx = (slot < nleft) ? leftovers[slot] : order[first_hit + (slot - nleft)];
                const SimplexRec *r = simplices + cur_sidx;
                const uint32_t pair = r->pair;
                iter = 0;
                out_slot = out_index[pair];
                if (out_slot >= hit_capacity)
                    ++n_dropped; // more GJK hits than contact records: the step reports PK_E_PAIR_OVERFLOW
                else if ((r->n & 0xFFu) != 4u)
                    fb = 1; // pad_simplex path (collision.cpp:191-248): rare, left to epa_kernel
                else
                {
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
                    {
                        ShapeView A = load_shape(bodies, ia);
                        ShapeView B = load_shape(bodies, ib);
                        es_put_shape(shm, 0, A, bodies);
                        es_put_shape(shm, 1, B, bodies);
                    }
                    const EpaInit *in = init + cur_sidx; // epa_init_kernel: planes and topology of the tetrahedron
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                    {
                        SupportPt s;
                        s.pa = d3{r->v[i][0], r->v[i][1], r->v[i][2]};
                        s.pb = d3{r->v[i][3], r->v[i][4], r->v[i][5]};
                        sl.set_vert(i, s, P(s));
                    }
                    if constexpr (!HEAP)
                    {
                        const int dirty = keys_dirty ? ES_KEYS : (hi < ES_KEYS ? hi : ES_KEYS);
                        for (int s = 4; s < dirty; ++s) shm.pop.th[t].key[s] = INF;
                        for (int s = ES_KEYS; s < gdirty; ++s) sl.gkey[s - ES_KEYS] = INF;
                        gdirty = ES_KEYS;
                        keys_dirty = false;
                    }
                    else
                        keys_dirty = true;
                    heap_size = 0;
                    const bool bad = in->bad != 0u;
#pragma unroll
                    for (int f = 0; f < 4; ++f)
                    {
                        const double2 *q = reinterpret_cast<const double2 *>(in->plane[f]);
                        const double2 n01 = q[0], n2d = q[1];
                        sl.store_plane(f, d3{n01.x, n01.y, n2d.x}, n2d.y);
                        sl.topo[f] = in->topo[f];
                        if constexpr (HEAP)
                        {
                            es_sift_up(hp, heap_size, n2d.y, static_cast<uint32_t>(f) | (static_cast<uint32_t>(f) << 8)); // push_face
                            ++heap_size;
                        }
                        else
                        {
                            shm.pop.th[t].key[f] = __double2float_rd(n2d.y);
                            shm.hz[f][t] = static_cast<uint32_t>(f) << 24; // push order of this batch (tie breaking)
                        }
                    }
                    fm0 = ~0xFull;
                    fm1 = ~0ull;
                    fm2 = (1ull << (ES_SLOTS - 128)) - 1ull;
                    hi = 4;
                    nfaces = 4;
                    nverts = 4;
                    stale_lb = 1e300;
                    batch_n = 4;
                    active = true;
                    if (bad) fb = 4;
                }
            }
        }

        // ---------------- S1: the sequential part of one iteration, one lane per pair ----------------
        int emit = 0; // horizon edges this lane hands to S2
        if (active && !fb)
        {
            do
            {
                // ---- pop_face (collision.cpp:397-408) ----
                int min_face = -1;
                if constexpr (HEAP)
                {
                    while (heap_size > 0) // skip obsolete entries: slot free, or re-used by a younger face
                    {
                        const uint32_t id = es_heap_pop(hp, heap_size);
                        const int f = static_cast<int>(id & 0xFFu);
                        if (is_free(f)) continue;
                        if (nfaces > ES_SLOTS && static_cast<uint32_t>(sl.topo[f] >> 48) != (id >> 8)) continue;
                        min_face = f;
                        break;
                    }
                }
                else
                {
                    float m = INF;
                    int cnt = 0, second = -1;
                    const int hs = hi < ES_KEYS ? hi : ES_KEYS;
                    const float4 *kq = reinterpret_cast<const float4 *>(shm.pop.th[t].key);
                    for (int s = 0; s < hs; s += 4) // slots in [hs, s+4) hold +inf: they never win and never count
                    {
                        const float4 k4 = kq[s >> 2];
                        const float kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                        {
                            if (kk[u] < m)
                            {
                                m = kk[u];
                                min_face = s + u;
                                cnt = 1;
                            }
                            else if (kk[u] == m)
                            {
                                if (cnt == 1) second = s + u;
                                ++cnt;
                            }
                        }
                    }
                    if (hi > ES_KEYS) // only polytopes past ≈45 iterations: same scan over the keys kept in the slab
                    {
                        const float4 *gq = reinterpret_cast<const float4 *>(sl.gkey);
                        for (int s = ES_KEYS; s < hi; s += 4)
                        {
                            const float4 k4 = gq[(s - ES_KEYS) >> 2];
                            const float kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                            {
                                if (kk[u] < m)
                                {
                                    m = kk[u];
                                    min_face = s + u;
                                    cnt = 1;
                                }
                                else if (kk[u] == m)
                                {
                                    if (cnt == 1) second = s + u;
                                    ++cnt;
                                }
                            }
                        }
                    }
                    if (min_face >= 0 && cnt > 1)
                    {
                        // several live faces share the minimal float key: compare the exact distances
                        double best = sl.plane[4 * min_face + 3];
                        bool tie = false;
                        if (cnt == 2)
                        {
                            const double d = sl.plane[4 * second + 3];
                            if (d < best)
                            {
                                best = d;
                                min_face = second;
                            }
                            else if (d == best)
                                tie = true;
                        }
                        else
                        {
                            for (int s = min_face + 1; s < hi; ++s)
                            {
                                if (key_of(s) != m) continue;
                                const double d = sl.plane[4 * s + 3];
                                if (d < best)
                                {
                                    best = d;
                                    min_face = s;
                                    tie = false;
                                }
                                else if (d == best)
                                    tie = true;
                            }
                        }
                        if (tie)
                        {
                            // the two cases that can be decided without the heap's history: see epa_scan_kernel
                            bool resolved = false;
                            if (best < stale_lb)
                            {
                                int n_old = 0, old_face = -1, first_new = -1;
                                for (int e = batch_n - 1; e >= 0; --e)
                                {
                                    const int s = hz_slot(shm.hz[e][t]);
                                    if (!is_free(s) && sl.plane[4 * s + 3] == best) first_new = s;
                                }
                                for (int s = 0; s < hi; ++s)
                                {
                                    if (key_of(s) != m || sl.plane[4 * s + 3] != best) continue; // (free slots have key +inf)
                                    bool in_batch = false;
                                    for (int e = 0; e < batch_n; ++e) in_batch = in_batch || hz_slot(shm.hz[e][t]) == s;
                                    if (!in_batch)
                                    {
                                        ++n_old;
                                        old_face = s;
                                    }
                                }
                                if (n_old == 1)
                                {
                                    min_face = old_face;
                                    resolved = true;
                                }
                                else if (n_old == 0 && first_new >= 0)
                                {
                                    min_face = first_new;
                                    resolved = true;
                                }
                            }
                            if (!resolved)
                            {
                                fb = 2;
                                break;
                            }
                        }
                    }
                }
                if (min_face < 0)
                {
                    valid[out_slot] = 0; // heap exhausted → nullopt (collision.cpp:459,502)
                    active = false;
                    break;
                }
                const double4 mf = sl.load_plane(min_face);
                const unsigned long long mt = sl.topo[min_face];
                if (mf.w > stale_lb) stale_lb = mf.w; // entries closer than the popped face have left the heap
                const bool finished = iter >= 64;     // best guess after the loop (collision.cpp:500-503)
                if (!finished) ++iter;
#pragma unroll
                for (int k = 0; k < 3; ++k) // the flood fill starts with these: on their way during the support evaluation
                {
                    const int b = es_adj(mt, k);
                    if (b != 0xFF)
                    {
                        es_prefetch(sl.plane + 4 * b);
                        es_prefetch(sl.topo + b);
                    }
                }
                const d3 mn{mf.x, mf.y, mf.z};
                SupportPt sp{}; // minkowski_support (collision.h:41-49)
                if (!finished)
                {
                    {
                        const ShapeView A = es_get_shape(shm, 0, bodies);
                        sp.pa = support(A, mn);
                    }
                    {
                        const ShapeView B = es_get_shape(shm, 1, bodies);
                        sp.pb = support(B, -mn);
                    }
                }
                const d3 p = P(sp);
                if (finished || dot(mn, p) - mf.w < 1e-6) // converged (collision.cpp:465-466)
                {
                    pending = true;
                    pend_face = min_face;
                    active = false;
                    break;
                }

                // ---- find_silhouette (collision.cpp:315-353): LIFO flood fill, edge order preserved ----
                bool bad = false;
                int nh = 0;
                {
                    kill_slot(min_face);
                    unsigned long long stack = 0;
                    int depth = 0;
                    unsigned long long cur = mt;
                    for (;;)
                    {
                        double4 nf[3];
                        unsigned long long nt[3];
                        bool live[3];
#pragma unroll
                        for (int i = 0; i < 3; ++i)
                        {
                            const int a = es_adj(cur, i);
                            live[i] = a != 0xFF && !is_free(a);
                            if (live[i])
                            {
                                nf[i] = sl.load_plane(a);
                                nt[i] = sl.topo[a];
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 3; ++i)
                        {
                            if (!live[i]) continue;
                            const int a = es_adj(cur, i);
                            if (is_free(a)) continue; // reached through two edges of this face: the first visit killed it
                            if (dot(d3{nf[i].x, nf[i].y, nf[i].z}, p) > nf[i].w + 1e-6)
                            {
                                kill_slot(a);
                                if (nf[i].w < stale_lb) stale_lb = nf[i].w; // its heap entry stays behind
                                if (depth < ES_STACK)
                                {
                                    stack = (stack << 8) | static_cast<unsigned long long>(a);
                                    ++depth;
#pragma unroll
                                    for (int k = 0; k < 3; ++k) // its neighbourhood will be needed when it is popped
                                    {
                                        const int b = es_adj(nt[i], k);
                                        if (b != 0xFF)
                                        {
                                            es_prefetch(sl.plane + 4 * b);
                                            es_prefetch(sl.topo + b);
                                        }
                                    }
                                }
                                else
                                    bad = true;
                            }
                            else
                            {
                                // horizon edge (cur.v[i], cur.v[i+1], a)
                                const int st = es_v(cur, i), en = es_v(cur, (i + 1) % 3);
                                if (nh < ES_HORIZON)
                                {
                                    shm.hz[nh][t] = static_cast<uint32_t>(st) | (static_cast<uint32_t>(en) << 7) | (static_cast<uint32_t>(a) << 14);
                                    ++nh;
                                }
                                else
                                    bad = true;
                            }
                        }
                        if (depth == 0) break;
                        cur = sl.topo[static_cast<int>(stack & 0xFFull)];
                        stack >>= 8;
                        --depth;
                    }
                }
                batch_n = 0;
                if (nh == 0 && !bad)
                {
                    iter = 64; // empty horizon → best remaining face (collision.cpp:469,500-503)
                    break;
                }
                const int nfree = __popcll(fm0) + __popcll(fm1) + __popcll(fm2);
                const bool full = nfree < nh || nverts >= ES_VERTS || (HEAP && heap_size + nh > ES_HEAP_MAX);
                if (bad || nh < 3 || full)
                {
                    fb = full ? 3 : 4;
                    break;
                }
                sl.set_vert(nverts, sp, p);
                shm.pnew[0][t] = p.x;
                shm.pnew[1][t] = p.y;
                shm.pnew[2][t] = p.z;
                shm.meta[t] = static_cast<uint32_t>(nverts) | (static_cast<uint32_t>(nfaces) << 8);
                shm.bad[t] = 0u;
                ++nverts;
                // slots of the new faces in horizon (= push) order, lowest free first
                for (int e = 0; e < nh; ++e)
                {
                    const int slot = take_slot();
                    if (slot >= hi) hi = slot + 1;
                    if (!HEAP && slot >= gdirty) gdirty = slot + 1;
                    ++nfaces;
                    shm.hz[e][t] |= static_cast<uint32_t>(slot) << 24;
                }
                emit = nh;
            } while (false);
        }

        // ---------------- S2: the new faces of all pairs of the warp, dealt out evenly over its lanes ----------------
        int incl = emit;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const int v = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += v;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        if (total > 0)
        {
            __syncwarp(); // horizons, new vertices and slots are in shared memory / the slabs
            for (int base = 0; base < total; base += 32)
            {
                const int i = base + lane;
                int lo = 0; // number of lanes whose inclusive count is ≤ i = the owner of edge i
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1)
                {
                    const int v = __shfl_sync(FULL, incl, lo + step - 1);
                    if (v <= i) lo += step;
                }
                const int L = lo & 31;
                const int o_incl = __shfl_sync(FULL, incl, L);
                const int o_cnt = __shfl_sync(FULL, emit, L);
                if (i < total)
                {
                    const EsSlab osl(wslab + static_cast<size_t>(L) * slab_bytes, HEAP);
                    ec_make_face<HEAP>(shm, osl, wbase + L, i - (o_incl - o_cnt), o_cnt);
                }
            }
            __syncwarp();
            if (emit)
            {
                if (shm.bad[t])
                    fb = 4;
                else
                {
                    if constexpr (HEAP)
                    {
                        const uint32_t serial0 = shm.meta[t] >> 8;
                        for (int e = 0; e < emit; ++e) // push_face in horizon order (collision.cpp:479)
                        {
                            const int slot = hz_slot(shm.hz[e][t]);
                            es_sift_up(hp, heap_size, sl.plane[4 * slot + 3], static_cast<uint32_t>(slot) | ((serial0 + static_cast<uint32_t>(e)) << 8));
                            ++heap_size;
                        }
                    }
                    batch_n = emit;
                }
            }
        }
    }
    if (n_valid) atomicAdd(counters + 0, n_valid);
    if (n_dropped) atomicAdd(counters + 1, n_dropped);
}

} // namespace pk
