// pk_solver.cuh — contact Jacobian set-up on the device: constraint_solver::setup_contacts
// (include/physkit/collision/constraint.h:1052-1104) = build_contact_jacobian (:874-953) +
// build_orthonormal_basis (:104-113) for every point of every manifold, the 2×2 friction block inverse and the
// warm-start impulses.  The serial Gauss-Seidel sweep that consumes the rows (:1107-1201) is not on this path.
//
// Input: the sorted manifold array of pk_manifold.cuh, the body state of pk_dynamics.cuh (velocities, inverse
// mass, world inverse inertia tensor) and the poses.  solver_valid_kernel (one thread per point slot, 4 per
// manifold) only decides which points yield a row (penetration > 0, constraint.h:893-894); a flag scan turns
// that into the row index and solver_compact_kernel into its inverse; solver_rows_kernel computes one row per
// thread and writes it in place — rows come out ordered by (pair key, point), with no staging copy of the
// 336-byte records.  FP64, no FMA, the reference's operation order: bit-identical to the oracle's restatement.
#pragma once

#include "pk_dynamics.cuh"
#include "pk_manifold.cuh"

namespace pk
{

struct SolverRow // jacobian_row (constraint.h:29-47), the fields setup_contacts fills
{
    double J_v[3], J_w_a[3], J_w_b[3];
    double M_eff, bias;
};
struct SolverPoint // contact_solver_point (constraint.h:1204-1214); mirrors pk_solver_point
{
    uint64_t key;
    uint32_t manifold; // index into the manifold array of this step (the `cache` pointer of the reference)
    uint32_t point;    // index of the contact in its manifold
    SolverRow normal, tangent1, tangent2;
    double friction_coeff, inv_m_11, inv_m_12, inv_m_22;
    double accumulated[3]; // warm start: normal, tangent 1, tangent 2 (constraint.h:1096-1098)
};
static_assert(sizeof(SolverRow) == 88 && sizeof(SolverPoint) == 16 + 3 * 88 + 7 * 8, "SolverPoint mirrors pk_solver_point");

struct SolverBody
{
    d3 pos, vel, ang_vel;
    dq q;
    double inv_mass, restitution, friction;
    dm3 iiw;
};
__device__ __forceinline__ SolverBody load_solver_body(const double *__restrict__ pos, const double *__restrict__ quat, const DynArrays &dy,
                                                       const double *__restrict__ material, uint32_t i, bool full)
{
    SolverBody b;
    b.pos = d3{pos[3ull * i], pos[3ull * i + 1], pos[3ull * i + 2]};
    b.q = dq{quat[4ull * i], quat[4ull * i + 1], quat[4ull * i + 2], quat[4ull * i + 3]};
    if (full)
    {
        b.vel = d3{dy.vel[3ull * i], dy.vel[3ull * i + 1], dy.vel[3ull * i + 2]};
        b.ang_vel = d3{dy.ang_vel[3ull * i], dy.ang_vel[3ull * i + 1], dy.ang_vel[3ull * i + 2]};
        b.inv_mass = dy.mass[2ull * i + 1];
        b.iiw = load_m3(dy.inertia_w + 18ull * i + 9);
        b.restitution = material[2ull * i];
        b.friction = material[2ull * i + 1];
    }
    return b;
}

struct ContactFrame
{
    d3 r_a, r_b, n;
    double penetration;
};
// constraint.h:883-893
__device__ __forceinline__ ContactFrame contact_frame(const SolverBody &a, const SolverBody &b, const ManifoldPoint &c)
{
    ContactFrame f;
    const d3 nn = normalized(mp_ld(c.normal));
    const d3 local_normal_b = rotate(conjugate(b.q), nn);
    f.r_a = rotate(a.q, mp_ld(c.local_a));
    f.r_b = rotate(b.q, mp_ld(c.local_b));
    f.n = rotate(b.q, local_normal_b);
    f.penetration = dot((b.pos + f.r_b) - (a.pos + f.r_a), f.n);
    return f;
}

__global__ void __launch_bounds__(256)
solver_valid_kernel(const ManifoldRec *__restrict__ man, uint64_t nman, const double *__restrict__ pos, const double *__restrict__ quat,
                    uint8_t *__restrict__ valid)
{
    const uint64_t slot = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (slot >= 4 * nman) return;
    const ManifoldRec *m = man + (slot >> 2);
    const uint32_t j = static_cast<uint32_t>(slot & 3u);
    uint8_t ok = 0;
    if (j < m->count)
    {
        const uint64_t key = m->key;
        DynArrays none{};
        const SolverBody a = load_solver_body(pos, quat, none, nullptr, static_cast<uint32_t>(key >> 32), false);
        const SolverBody b = load_solver_body(pos, quat, none, nullptr, static_cast<uint32_t>(key & 0xFFFFFFFFu), false);
        ok = contact_frame(a, b, m->pt[j]).penetration > 0.0 ? 1 : 0; // `<= 0 → nullopt`, constraint.h:894
    }
    valid[slot] = ok;
}

__device__ __forceinline__ void st3(double *p, d3 v)
{
    p[0] = v.x;
    p[1] = v.y;
    p[2] = v.z;
}

// slot of every row (the inverse of the flag scan), so that solver_rows_kernel runs one thread per ROW: most
// manifolds hold one or two points, a thread per point slot would leave three lanes of four idle in the
// 150-register kernel (ncu r1b: 7 of 32 lanes active).
__global__ void __launch_bounds__(256)
solver_compact_kernel(const uint8_t *__restrict__ valid, const uint32_t *__restrict__ row_index, uint64_t nslots,
                      uint32_t *__restrict__ slot_of_row)
{
    const uint64_t slot = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (slot < nslots && valid[slot]) slot_of_row[row_index[slot]] = static_cast<uint32_t>(slot);
}

__global__ void __launch_bounds__(128)
solver_rows_kernel(const ManifoldRec *__restrict__ man, const unsigned long long *__restrict__ nrows_ptr, const double *__restrict__ pos,
                   const double *__restrict__ quat, DynArrays dy, const double *__restrict__ material,
                   const uint32_t *__restrict__ slot_of_row, double dt, double gravity_norm, SolverPoint *__restrict__ out)
{
    const uint64_t row = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (row >= *nrows_ptr) return;
    const uint64_t slot = slot_of_row[row];
    const ManifoldRec *m = man + (slot >> 2);
    const uint32_t j = static_cast<uint32_t>(slot & 3u);
    const uint64_t key = m->key;
    const SolverBody a = load_solver_body(pos, quat, dy, material, static_cast<uint32_t>(key >> 32), true);
    const SolverBody b = load_solver_body(pos, quat, dy, material, static_cast<uint32_t>(key & 0xFFFFFFFFu), true);
    const ManifoldPoint c = m->pt[j];
    const ContactFrame f = contact_frame(a, b, c);
    constexpr double bias_factor = .1, linear_slop = 0.005; // constraint.h:880-881
    const double restitution_threshold = 2 * gravity_norm * dt; // :1071

    SolverPoint p;
    p.key = key;
    p.manifold = static_cast<uint32_t>(slot >> 2);
    p.point = j;
    auto fill = [&](SolverRow &row, d3 dir)
    {
        const d3 jwa = cross(f.r_a, dir), jwb = -cross(f.r_b, dir);
        const d3 k_a = mul(a.iiw, jwa), k_b = mul(b.iiw, jwb);
        st3(row.J_v, dir);
        st3(row.J_w_a, jwa);
        st3(row.J_w_b, jwb);
        row.M_eff = 1.0 / (((a.inv_mass + b.inv_mass) + dot(jwa, k_a)) + dot(jwb, k_b));
    };
    // normal row (:897-930)
    fill(p.normal, f.n);
    {
        const d3 v_ca = a.vel + cross(a.ang_vel, f.r_a);
        const d3 v_cb = b.vel + cross(b.ang_vel, f.r_b);
        const double v_rel_n = dot(v_ca - v_cb, f.n);
        const double restitution_coeff = dmax(a.restitution, b.restitution);
        double restitution_bias = 0.0;
        if (v_rel_n < -restitution_threshold) restitution_bias = restitution_coeff * v_rel_n;
        const double baumgarte_bias = (-bias_factor / dt) * dmax(0.0, f.penetration - linear_slop);
        p.normal.bias = dmin(baumgarte_bias, restitution_bias);
    }
    // friction rows (:933-957); build_orthonormal_basis (:104-113)
    const double sign = copysign(1.0, f.n.z);
    const double aa = -1.0 / (sign + f.n.z);
    const double bb = f.n.x * f.n.y * aa;
    const d3 t1{1.0 + sign * f.n.x * f.n.x * aa, sign * bb, -sign * f.n.x};
    const d3 t2{bb, sign + f.n.y * f.n.y * aa, -f.n.y};
    fill(p.tangent1, t1);
    fill(p.tangent2, t2);
    p.tangent1.bias = 0.0;
    p.tangent2.bias = 0.0;
    // setup_contacts (:1075-1098)
    p.friction_coeff = sqrt(a.friction * b.friction);
    const double m11 = 1.0 / p.tangent1.M_eff, m22 = 1.0 / p.tangent2.M_eff;
    const double m12 = dot(mp_ld(p.tangent1.J_w_a), mul(a.iiw, mp_ld(p.tangent2.J_w_a))) +
                       dot(mp_ld(p.tangent1.J_w_b), mul(b.iiw, mp_ld(p.tangent2.J_w_b)));
    const double det = m11 * m22 - m12 * m12;
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
    p.accumulated[0] = c.normal_impulse;
    p.accumulated[1] = c.tangent_impulses[0];
    p.accumulated[2] = c.tangent_impulses[1];
    out[row] = p;
}

} // namespace pk
