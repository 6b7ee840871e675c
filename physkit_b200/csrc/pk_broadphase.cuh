// pk_broadphase.cuh — per-step LBVH broadphase.
//
// Reproduces the RESULT of broad_phase::update_node + calculate_pairs
// (reference collision_phases.h:371-436, src/bvh.cpp:475-514) with a stateless rebuild:
//
//   K1  bounds_fat_kernel   mesh::instance::bounds() (src/mesh.cpp:398 → bounds.h:142-158) and the
//                           fat-AABB rule of dynamic_bvh::update_leaf (src/bvh.cpp:483-508); FP64,
//                           bit-identical.  Also reduces the scene bounds of the box centres.
//   K1b morton_kernel       30-bit Morton key of the centre (+ world tiling), float boxes rounded outward
//   K2  radix sort          (pk_sort.cuh) keys → Morton order
//   K3  leaf_kernel         gathers the sorted leaves: 64-byte exact record + 32-byte float node
//   K4  hierarchy_kernel    Apetrei-style bottom-up agglomeration: hierarchy emission and refit in
//                           one pass, one atomic exchange per internal node
//   K4b rope_kernel         escape pointers for stackless traversal
//   K5  overlap_kernel      one query per lane, lanes in Morton order; each leaf starts at its own rope
//                           so every unordered pair is met exactly once (j > i in leaf order); float
//                           prefilter on 32-byte nodes, exact inclusive FP64 test (bounds.h:87-92) on
//                           the 64-byte leaf record, warp-aggregated append of u64 pair keys
//   K6  radix sort          pair keys ascending = the reference's set as a sorted (i,j) list
//
// The tree's shape is irrelevant to the result: the pair set is a pure function of the stored
// boxes, aabb::intersects and the move history (proven against the faithful
// incremental implementation in tests/test_oracle_bvh_shapes.py).
#pragma once

#include "pk_common.cuh"
#include "pk_sort.cuh"

namespace pk
{

constexpr uint32_t NODE_SENTINEL = 0xFFFFFFFFu;

// 32-byte traversal node.  Internal: left = child node, leaves: left = body id.
struct alignas(16) NodeF
{
    float lo[3];
    float hi[3];
    uint32_t left;
    uint32_t rope;
};
static_assert(sizeof(NodeF) == 32, "NodeF is one sector");

// 64-byte exact leaf record (sorted order).
struct alignas(16) LeafRec
{
    double box[6]; // stored (fat) box: min xyz, max xyz
    uint32_t id;
    int32_t last_move; // epoch of the last re-insertion, -1 = never
    int32_t create;    // epoch at which the body was created
    uint32_t world;
};
static_assert(sizeof(LeafRec) == 64, "LeafRec is two sectors");

struct BodyState
{
    double *stored;     // [n][6]
    int32_t *last_move; // [n]
    int32_t *create;    // [n]
    uint8_t *alive;     // [n] alive in the previous step
};

// order-preserving float <-> uint for atomicMin/atomicMax
__device__ __forceinline__ uint32_t f2ord(float f)
{
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// scene[0..2] = min centre (ordered uint), scene[3..5] = max centre
// scene == nullptr: counters only (pk_gjk_epa_batch must leave the last step's scene box to pk_raycast)
__global__ void scene_reset_kernel(uint32_t *scene, unsigned long long *counters, int ncounters)
{
    int t = threadIdx.x;
    if (scene)
    {
        if (t < 3) scene[t] = 0xFFFFFFFFu;
        else if (t < 6) scene[t] = 0u;
    }
    if (t < ncounters) counters[t] = 0ull;
}

// ---------------------------------------------------------------------------------------------
// K1: true bounds + fat rule.  One thread per body.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bounds_fat_kernel(const ShapeRec *__restrict__ shapes, const double *__restrict__ pos, const double *__restrict__ quat,
                  const double *__restrict__ disp, const uint32_t *__restrict__ shape_id,
                  const uint8_t *__restrict__ flags, uint32_t n, int mode_query, int32_t epoch, BodyState st,
                  uint32_t *__restrict__ scene, unsigned long long *__restrict__ moved_counter)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float cmin[3] = {3.0e38f, 3.0e38f, 3.0e38f}, cmax[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    bool moved = false;
    if (i < n)
    {
        uint8_t fl = flags[i];
        bool alive = (fl & FLAG_ALIVE) != 0;
        bool was = st.alive[i] != 0;
        if (!alive)
        {
            if (was)
            {
                st.alive[i] = 0;
                st.last_move[i] = -1;
            }
        }
        else
        {
            const ShapeRec *s = shapes + shape_id[i];
            const double2 *sp = reinterpret_cast<const double2 *>(s);
            double2 s0 = __ldg(sp), s1 = __ldg(sp + 1), s2 = __ldg(sp + 2);
            int kind = __ldg(reinterpret_cast<const int *>(sp + 3));
            d3 lmin, lmax;
            if (kind == KIND_OBB)
            {
                lmax = {s0.x, s0.y, s1.x};
                lmin = -lmax;
            }
            else if (kind == KIND_SPHERE)
            {
                lmax = {s0.x, s0.x, s0.x};
                lmin = -lmax;
            }
            else
            {
                lmin = {s0.x, s0.y, s1.x};
                lmax = {s1.y, s2.x, s2.y};
            }
            d3 tmin, tmax;
            if (kind == KIND_AABB)
            {
                tmin = lmin;
                tmax = lmax;
            }
            else
            {
                // aabb::operator*(unit_quat) (bounds.h:142-158): corner 0 seeds min and max,
                // corner i: bit0→x, bit1→y, bit2→z, set ⇒ max (bounds.h:65-71); then + pos (:112-113)
                const double2 *qp = reinterpret_cast<const double2 *>(quat + 4ull * i);
                double2 q0 = __ldg(qp), q1 = __ldg(qp + 1);
                dq q{q0.x, q0.y, q1.x, q1.y};
                d3 p0 = rotate(q, lmin);
                tmin = p0;
                tmax = p0;
#pragma unroll
                for (int c = 1; c < 8; ++c)
                {
                    d3 corner{(c & 1) ? lmax.x : lmin.x, (c & 2) ? lmax.y : lmin.y, (c & 4) ? lmax.z : lmin.z};
                    d3 pt = rotate(q, corner);
                    tmin.x = dmin(tmin.x, pt.x);
                    tmin.y = dmin(tmin.y, pt.y);
                    tmin.z = dmin(tmin.z, pt.z);
                    tmax.x = dmax(tmax.x, pt.x);
                    tmax.y = dmax(tmax.y, pt.y);
                    tmax.z = dmax(tmax.z, pt.z);
                }
                d3 pp{pos[3ull * i], pos[3ull * i + 1], pos[3ull * i + 2]};
                tmin = tmin + pp;
                tmax = tmax + pp;
            }
            double *sb = st.stored + 6ull * i;
            d3 smin, smax;
            if (mode_query || !was)
            {
                // dynamic_bvh::add stores the exact box (bvh.h:294-302); not marked moved
                smin = tmin;
                smax = tmax;
                if (!was)
                {
                    st.alive[i] = 1;
                    st.create[i] = epoch;
                    st.last_move[i] = -1;
                }
                sb[0] = smin.x; sb[1] = smin.y; sb[2] = smin.z;
                sb[3] = smax.x; sb[4] = smax.y; sb[5] = smax.z;
            }
            else
            {
                smin = {sb[0], sb[1], sb[2]};
                smax = {sb[3], sb[4], sb[5]};
                bool is_static = (fl & FLAG_STATIC) != 0;
                // aabb::contains(aabb) (bounds.h:80-85), inclusive
                bool contains = (tmin.x >= smin.x && tmax.x <= smax.x) && (tmin.y >= smin.y && tmax.y <= smax.y) &&
                                (tmin.z >= smin.z && tmax.z <= smax.z);
                if (!is_static && !contains)
                {
                    // src/bvh.cpp:487-506
                    const double margin = .1;
                    smin = {tmin.x - margin, tmin.y - margin, tmin.z - margin};
                    smax = {tmax.x + margin, tmax.y + margin, tmax.z + margin};
                    d3 dd{disp[3ull * i], disp[3ull * i + 1], disp[3ull * i + 2]};
                    if (dd.x < 0.0) smin.x = smin.x + dd.x; else smax.x = smax.x + dd.x;
                    if (dd.y < 0.0) smin.y = smin.y + dd.y; else smax.y = smax.y + dd.y;
                    if (dd.z < 0.0) smin.z = smin.z + dd.z; else smax.z = smax.z + dd.z;
                    sb[0] = smin.x; sb[1] = smin.y; sb[2] = smin.z;
                    sb[3] = smax.x; sb[4] = smax.y; sb[5] = smax.z;
                    st.last_move[i] = epoch;
                    moved = true;
                }
            }
            float cx = static_cast<float>(0.5 * (smin.x + smax.x));
            float cy = static_cast<float>(0.5 * (smin.y + smax.y));
            float cz = static_cast<float>(0.5 * (smin.z + smax.z));
            cmin[0] = cmax[0] = cx;
            cmin[1] = cmax[1] = cy;
            cmin[2] = cmax[2] = cz;
        }
    }
    // block reduction of centre bounds + moved count
    unsigned m = __ballot_sync(0xFFFFFFFFu, moved);
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            cmin[k] = fminf(cmin[k], __shfl_xor_sync(0xFFFFFFFFu, cmin[k], o));
            cmax[k] = fmaxf(cmax[k], __shfl_xor_sync(0xFFFFFFFFu, cmax[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0)
    {
        if (cmin[0] <= cmax[0])
        {
#pragma unroll
            for (int k = 0; k < 3; ++k)
            {
                atomicMin(scene + k, f2ord(cmin[k]));
                atomicMax(scene + 3 + k, f2ord(cmax[k]));
            }
        }
        if (m) atomicAdd(moved_counter, static_cast<unsigned long long>(__popc(m)));
    }
}

__device__ __forceinline__ uint32_t expand_bits10(uint32_t v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

// World tiling: world w is shifted by (wx, wy, wz)·tile in FLOAT space only (Morton keys and the
// conservative float boxes); the exact FP64 test never sees the shift.  Worlds therefore separate
// spatially in the tree and a query never descends into another world's subtree.
struct WorldTiling
{
    uint32_t num_worlds;
    uint32_t grid; // worlds per axis (ceil(cbrt(num_worlds)))
};

__device__ __forceinline__ void world_offset(const WorldTiling &wt, uint32_t w, float tile, float off[3])
{
    if (wt.num_worlds <= 1)
    {
        off[0] = off[1] = off[2] = 0.f;
        return;
    }
    uint32_t g = wt.grid;
    off[0] = static_cast<float>(w % g) * tile;
    off[1] = static_cast<float>((w / g) % g) * tile;
    off[2] = static_cast<float>(w / (g * g)) * tile;
}

__device__ __forceinline__ float scene_tile(const uint32_t *scene)
{
    float ext = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) ext = fmaxf(ext, ord2f(scene[3 + k]) - ord2f(scene[k]));
    return ext * 1.25f + 8.0f; // > extent of any world incl. the largest fat box half-size margin
}

// ---------------------------------------------------------------------------------------------
// K1b: Morton keys.  Dead bodies get the all-ones key and sort to the end.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
morton_kernel(const double *__restrict__ stored, const uint8_t *__restrict__ alive, const uint32_t *__restrict__ world_id,
              uint32_t n, const uint32_t *__restrict__ scene, WorldTiling wt, uint64_t *__restrict__ keys,
              uint32_t *__restrict__ vals)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    vals[i] = i;
    if (!alive[i])
    {
        keys[i] = 0xFFFFFFFFFFFFFFFFull;
        return;
    }
    float tile = scene_tile(scene);
    float off[3];
    world_offset(wt, world_id ? world_id[i] : 0u, tile, off);
    float span_w = (wt.num_worlds > 1) ? tile * static_cast<float>(wt.grid) : 0.f;
    const double *sb = stored + 6ull * i;
    uint32_t q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        float lo = ord2f(scene[k]);
        float span = fmaxf(ord2f(scene[3 + k]) - lo + span_w, 1e-30f);
        float c = static_cast<float>(0.5 * (sb[k] + sb[3 + k])) + off[k];
        float t = (c - lo) / span;
        t = fminf(fmaxf(t * 1024.f, 0.f), 1023.f);
        q[k] = static_cast<uint32_t>(t);
    }
    uint32_t code = (expand_bits10(q[0]) << 2) | (expand_bits10(q[1]) << 1) | expand_bits10(q[2]);
    keys[i] = static_cast<uint64_t>(code);
}

// ---------------------------------------------------------------------------------------------
// K3: gather sorted leaves.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
leaf_kernel(const uint32_t *__restrict__ sorted_ids, uint32_t m, const double *__restrict__ stored,
            const int32_t *__restrict__ last_move, const int32_t *__restrict__ create,
            const uint32_t *__restrict__ world_id, const uint32_t *__restrict__ scene, WorldTiling wt,
            LeafRec *__restrict__ leaves, NodeF *__restrict__ nodes, int32_t *__restrict__ merge_flag)
{
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    uint32_t id = sorted_ids[p];
    const double2 *sb = reinterpret_cast<const double2 *>(stored + 6ull * id);
    double2 b0 = sb[0], b1 = sb[1], b2 = sb[2];
    uint32_t w = world_id ? world_id[id] : 0u;
    LeafRec r;
    r.box[0] = b0.x; r.box[1] = b0.y; r.box[2] = b1.x;
    r.box[3] = b1.y; r.box[4] = b2.x; r.box[5] = b2.y;
    r.id = id;
    r.last_move = last_move[id];
    r.create = create[id];
    r.world = w;
    double2 *dst = reinterpret_cast<double2 *>(leaves + p);
    const double2 *src = reinterpret_cast<const double2 *>(&r);
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];

    float off[3];
    world_offset(wt, w, scene_tile(scene), off);
    NodeF nd;
    // outward rounding keeps the float box a superset of the double box even after the shift
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        nd.lo[k] = __fadd_rd(__double2float_rd(r.box[k]), off[k]);
        nd.hi[k] = __fadd_ru(__double2float_ru(r.box[3 + k]), off[k]);
    }
    nd.left = id;
    nd.rope = NODE_SENTINEL;
    float4 *nd4 = reinterpret_cast<float4 *>(nodes + (m - 1) + p);
    const float4 *ns = reinterpret_cast<const float4 *>(&nd);
    nd4[0] = ns[0];
    nd4[1] = ns[1];
    if (p + 1 < m) merge_flag[p] = -1;
}

// δ(i) between sorted leaves i and i+1; larger = split nearer the root.  Equal keys fall back to
// the index so the hierarchy is always a proper binary tree (Karras' tie-break).
struct Delta
{
    uint64_t hi;
    uint32_t lo;
};
__device__ __forceinline__ Delta delta_at(const uint64_t *__restrict__ keys, uint32_t i)
{
    Delta d;
    d.hi = keys[i] ^ keys[i + 1];
    d.lo = i ^ (i + 1);
    return d;
}
__device__ __forceinline__ bool delta_less(const Delta &a, const Delta &b)
{
    return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo);
}

// ---------------------------------------------------------------------------------------------
// K4: bottom-up hierarchy + refit.  Internal node index = split position (node s separates leaf
// s from leaf s+1), so a subtree covering leaves [l, r] hangs under node r (as its left child) or
// node l-1 (as its right child), whichever split is lower in the tree.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
hierarchy_kernel(const uint64_t *__restrict__ keys, uint32_t m, NodeF *nodes, uint32_t *right_child,
                 uint32_t *range_last, int32_t *merge_flag, uint32_t *root_out)
{
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    uint32_t l = p, r = p;
    uint32_t cur = (m - 1) + p;
    const float4 *me = reinterpret_cast<const float4 *>(nodes + cur);
    float4 a0 = me[0], a1 = me[1];
    float lo[3] = {a0.x, a0.y, a0.z};
    float hi[3] = {a0.w, a1.x, a1.y};
    for (;;)
    {
        if (l == 0 && r == m - 1)
        {
            *root_out = cur;
            return;
        }
        bool as_left;
        if (l == 0)
            as_left = true;
        else if (r == m - 1)
            as_left = false;
        else
            as_left = delta_less(delta_at(keys, r), delta_at(keys, l - 1));
        uint32_t parent = as_left ? r : l - 1;
        if (as_left)
            nodes[parent].left = cur;
        else
            right_child[parent] = cur;
        __threadfence();
        int32_t other = atomicExch(merge_flag + parent, static_cast<int32_t>(as_left ? l : r));
        if (other == -1) return; // first to arrive: the sibling's thread finishes this node
        __threadfence();
        uint32_t sib;
        if (as_left)
        {
            r = static_cast<uint32_t>(other);
            sib = __ldcg(right_child + parent);
        }
        else
        {
            l = static_cast<uint32_t>(other);
            sib = __ldcg(&nodes[parent].left);
        }
        const float4 *sp = reinterpret_cast<const float4 *>(nodes + sib);
        float4 s0 = __ldcg(sp), s1 = __ldcg(sp + 1);
        lo[0] = fminf(lo[0], s0.x); lo[1] = fminf(lo[1], s0.y); lo[2] = fminf(lo[2], s0.z);
        hi[0] = fmaxf(hi[0], s0.w); hi[1] = fmaxf(hi[1], s1.x); hi[2] = fmaxf(hi[2], s1.y);
        uint32_t left = as_left ? cur : sib;
        float4 *dst = reinterpret_cast<float4 *>(nodes + parent);
        dst[0] = make_float4(lo[0], lo[1], lo[2], hi[0]);
        dst[1] = make_float4(hi[1], hi[2], __uint_as_float(left), __uint_as_float(NODE_SENTINEL));
        range_last[parent] = r;
        cur = parent;
    }
}

// K4b: rope(node covering [.., last]) = right child of split `last` (next subtree in DFS order).
__global__ void __launch_bounds__(256)
rope_kernel(uint32_t m, NodeF *nodes, const uint32_t *__restrict__ right_child, const uint32_t *__restrict__ range_last)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * m - 1) return;
    uint32_t last = (i < m - 1) ? range_last[i] : i - (m - 1);
    nodes[i].rope = (last == m - 1) ? NODE_SENTINEL : right_child[last];
}

// ---------------------------------------------------------------------------------------------
// K5: self-overlap traversal.
// ---------------------------------------------------------------------------------------------
constexpr int OVERLAP_THREADS = 128;
constexpr int OVERLAP_STAGE = 128; // pair keys staged per warp before one reservation in the global list

// Where the pairs go.  LIST: appended to out_keys in the order they are found; a radix sort brings them into key order
// (K6).  ROWS: the pair (lo id, hi id) is dropped into the row of its lo id — row_count[lo] is bumped, rows[lo][slot] =
// hi — and pair_rows_emit_kernel writes the rows out one after the other, each sorted: key order without a sort of the
// whole list, because a body has a few dozen partners at most.  A row that overflows (PAIR_ROW partners with a larger
// id) is counted in row_overflow; the host then repeats the step in LIST form.
constexpr uint32_t PAIR_ROW = 64;
struct PairRows
{
    uint32_t *row_count;              // [n], zeroed before the launch
    uint32_t *rows;                   // [n][PAIR_ROW]
    unsigned long long *row_overflow; // pairs that found their row full
};

template <bool ROWS>
__device__ __forceinline__ void overlap_flush(const uint64_t *my_stage, int fill, int lane, uint64_t *__restrict__ out_keys, uint64_t capacity,
                                              unsigned long long *__restrict__ pair_counter, const PairRows &pr)
{
    constexpr unsigned FULL = 0xFFFFFFFFu;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(pair_counter, static_cast<unsigned long long>(fill));
    if constexpr (ROWS)
    {
        unsigned over = 0;
        for (int i = lane; i < fill; i += 32)
        {
            const uint64_t key = my_stage[i];
            const uint32_t lo = static_cast<uint32_t>(key >> 32);
            const uint32_t slot = atomicAdd(pr.row_count + lo, 1u);
            if (slot < PAIR_ROW)
                pr.rows[static_cast<uint64_t>(lo) * PAIR_ROW + slot] = static_cast<uint32_t>(key);
            else
                ++over;
        }
        if (__any_sync(FULL, over != 0u) && over) atomicAdd(pr.row_overflow, static_cast<unsigned long long>(over));
    }
    else
    {
        base = __shfl_sync(FULL, base, 0);
        for (int i = lane; i < fill; i += 32)
            if (base + i < capacity) out_keys[base + i] = my_stage[i];
    }
}

template <bool ROWS>
__global__ void __launch_bounds__(OVERLAP_THREADS)
overlap_kernel(const NodeF *__restrict__ nodes, const LeafRec *__restrict__ leaves, uint32_t m, uint32_t p_begin,
               uint32_t p_end, int mode_query, uint64_t *__restrict__ out_keys, uint64_t capacity,
               unsigned long long *__restrict__ pair_counter, PairRows pr)
{
    // Emission: a single global counter bumped once per pair was 71 % of this kernel's stall samples
    // (ncu r1).  Pairs are staged per warp in shared memory and the list is reserved in chunks.
    __shared__ uint64_t stage[OVERLAP_THREADS / 32][OVERLAP_STAGE];
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t *my_stage = stage[warp];
    int fill = 0; // warp-uniform

    const uint32_t p = p_begin + blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = p < p_end;
    double mnx = 0, mny = 0, mnz = 0, mxx = 0, mxy = 0, mxz = 0;
    float qlx = 0, qly = 0, qlz = 0, qhx = 0, qhy = 0, qhz = 0;
    uint32_t my_id = 0, my_world = 0;
    int32_t my_move = 0, my_create = 0;
    uint32_t node = NODE_SENTINEL;
    if (live)
    {
        const double2 *lp = reinterpret_cast<const double2 *>(leaves + p);
        double2 b0 = __ldg(lp), b1 = __ldg(lp + 1), b2 = __ldg(lp + 2);
        int4 meta = __ldg(reinterpret_cast<const int4 *>(lp + 3));
        mnx = b0.x; mny = b0.y; mnz = b1.x; mxx = b1.y; mxy = b2.x; mxz = b2.y;
        my_id = static_cast<uint32_t>(meta.x);
        my_move = meta.y;
        my_create = meta.z;
        my_world = static_cast<uint32_t>(meta.w);
        const float4 *me = reinterpret_cast<const float4 *>(nodes + (m - 1) + p);
        float4 f0 = __ldg(me), f1 = __ldg(me + 1);
        qlx = f0.x; qly = f0.y; qlz = f0.z; qhx = f0.w; qhy = f1.x; qhz = f1.y;
        node = __float_as_uint(f1.w); // own rope: everything to the right in DFS order
    }
    const uint32_t first_leaf = m - 1;
    while (__any_sync(FULL, node != NODE_SENTINEL))
    {
        uint64_t key = 0;
        bool emit = false;
        if (node != NODE_SENTINEL)
        {
            const float4 *np = reinterpret_cast<const float4 *>(nodes + node);
            float4 n0 = __ldg(np), n1 = __ldg(np + 1);
            bool hit = (qlx <= n0.w && qhx >= n0.x) && (qly <= n1.x && qhy >= n0.y) && (qlz <= n1.y && qhz >= n0.z);
            uint32_t next = __float_as_uint(n1.w);
            if (hit)
            {
                if (node >= first_leaf)
                {
                    const double2 *op = reinterpret_cast<const double2 *>(leaves + (node - first_leaf));
                    double2 o0 = __ldg(op), o1 = __ldg(op + 1), o2 = __ldg(op + 2);
                    int4 om = __ldg(reinterpret_cast<const int4 *>(op + 3));
                    // aabb::intersects (bounds.h:87-92), inclusive on all six comparisons
                    bool ov = (mnx <= o1.y && mxx >= o0.x) && (mny <= o2.x && mxy >= o0.y) && (mnz <= o2.y && mxz >= o1.x);
                    bool pass = mode_query ? true : (my_move >= om.z || om.y >= my_create);
                    if (ov && pass && static_cast<uint32_t>(om.w) == my_world)
                    {
                        uint32_t oid = static_cast<uint32_t>(om.x);
                        uint32_t lo_id = my_id < oid ? my_id : oid, hi_id = my_id < oid ? oid : my_id;
                        key = (static_cast<uint64_t>(lo_id) << 32) | hi_id;
                        emit = true;
                    }
                }
                else
                    next = __float_as_uint(n1.z); // descend to the left child
            }
            node = next;
        }
        const unsigned em = __ballot_sync(FULL, emit);
        if (em)
        {
            if (emit) my_stage[fill + __popc(em & ((1u << lane) - 1u))] = key;
            fill += __popc(em);
            if (fill > OVERLAP_STAGE - 32)
            {
                __syncwarp();
                overlap_flush<ROWS>(my_stage, fill, lane, out_keys, capacity, pair_counter, pr);
                __syncwarp();
                fill = 0;
            }
        }
    }
    if (fill)
    {
        __syncwarp();
        overlap_flush<ROWS>(my_stage, fill, lane, out_keys, capacity, pair_counter, pr);
    }
}

// ---------------------------------------------------------------------------------------------
// K6 (ROWS): the rows written out in id order, each sorted — the pair keys in ascending order.
//   pair_rows_sum_kernel   partners per tile of 256 bodies   → tile_sum_scan_kernel (pk_sort.cuh)
//   pair_rows_emit_kernel  one thread per body: its row into local memory, insertion sort (a row holds 14 ids on
//                          average in BASELINE C3), keys (id << 32 | partner) from the body's offset on
// ---------------------------------------------------------------------------------------------
constexpr int ROWS_TILE = 256;
static_assert(ROWS_TILE == SORT_THREADS, "block_exclusive_scan_256");

__global__ void __launch_bounds__(ROWS_TILE)
pair_rows_sum_kernel(const uint32_t *__restrict__ row_count, uint32_t n, uint32_t *__restrict__ tile_sum)
{
    __shared__ uint32_t warp_sums[ROWS_TILE / 32];
    const uint32_t i = blockIdx.x * ROWS_TILE + threadIdx.x;
    uint32_t c = i < n ? row_count[i] : 0u;
    c = c < PAIR_ROW ? c : PAIR_ROW;
    uint32_t tot;
    block_exclusive_scan_256(c, warp_sums, &tot);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(ROWS_TILE)
pair_rows_emit_kernel(const uint32_t *__restrict__ row_count, const uint32_t *__restrict__ rows, uint32_t n,
                      const uint32_t *__restrict__ tile_sum /* scanned */, uint64_t *__restrict__ out_keys, uint64_t capacity)
{
    __shared__ uint32_t warp_sums[ROWS_TILE / 32];
    const uint32_t i = blockIdx.x * ROWS_TILE + threadIdx.x;
    uint32_t c = i < n ? row_count[i] : 0u;
    c = c < PAIR_ROW ? c : PAIR_ROW;
    uint32_t tot;
    const uint64_t off = static_cast<uint64_t>(tile_sum[blockIdx.x]) + block_exclusive_scan_256(c, warp_sums, &tot);
    if (c == 0u) return;
    uint32_t v[PAIR_ROW];
    const uint4 *row = reinterpret_cast<const uint4 *>(rows + static_cast<uint64_t>(i) * PAIR_ROW);
    for (uint32_t k = 0; k < c; k += 4)
    {
        const uint4 q = __ldcs(row + (k >> 2));
        v[k] = q.x;
        v[k + 1] = q.y;
        v[k + 2] = q.z;
        v[k + 3] = q.w;
    }
    for (uint32_t k = 1; k < c; ++k)
    {
        const uint32_t x = v[k];
        uint32_t j = k;
        while (j > 0 && v[j - 1] > x)
        {
            v[j] = v[j - 1];
            --j;
        }
        v[j] = x;
    }
    const uint64_t hi = static_cast<uint64_t>(i) << 32;
    for (uint32_t k = 0; k < c; ++k)
        if (off + k < capacity) out_keys[off + k] = hi | v[k];
}

// Static bodies never query and are never updated, so two static bodies can never pair
// (src/world.cpp:24, collision_phases.h:405-433).  In the stateless rule this falls out of
// last_move = -1 < create for both, so no extra test is needed in overlap_kernel.

} // namespace pk
