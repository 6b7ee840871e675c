// pk_dynamics.cuh — rigid-body state on the device and the two per-body loops of world::step_impl
// (SURVEY §8 f3): semi_implicit_euler (include/physkit/detail/integrate.h:17-47), particle::apply_force /
// angular_accel / clear_forces / update_derived_state (include/physkit/core/particle.h:72-105, 140-146),
// src/world.cpp:22-34 (loop A) and :50-55 (loop B).
//
// With these two kernels poses never leave the device between steps: loop A leaves vel·dt in the
// displacement array that bounds_fat_kernel reads (the `vel * dt` argument of broad_phase::update_node),
// loop B moves pos / quat in place.  One thread per body, FP64, -fmad=false like everything else.
// Eigen's fixed-size 3×3 kernels are restated as coefficient sums in k order, (k0 + k1) + k2, like dot();
// the quaternion product order and sin / cos are unpinned at the ulp level (see oracle/pk_oracle.hpp).
#pragma once

#include "pk_common.cuh"

namespace pk
{

struct dm3
{
    double m[3][3]; // [row][col]
};
__device__ __forceinline__ dm3 mul(const dm3 &a, const dm3 &b)
{
    dm3 c;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c.m[i][j] = (a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j]) + a.m[i][2] * b.m[2][j];
    return c;
}
__device__ __forceinline__ d3 mul(const dm3 &a, d3 v)
{
    return {(a.m[0][0] * v.x + a.m[0][1] * v.y) + a.m[0][2] * v.z, (a.m[1][0] * v.x + a.m[1][1] * v.y) + a.m[1][2] * v.z,
            (a.m[2][0] * v.x + a.m[2][1] * v.y) + a.m[2][2] * v.z};
}
__device__ __forceinline__ dm3 transpose(const dm3 &a)
{
    dm3 t;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) t.m[i][j] = a.m[j][i];
    return t;
}
// Eigen QuaternionBase::toRotationMatrix (lin_alg.h:530-535)
__device__ __forceinline__ dm3 to_rotation_matrix(dq q)
{
    const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    dm3 r;
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
__device__ __forceinline__ dq qmul(dq a, dq b) // lin_alg.h:478-483
{
    return {((a.w * b.x + a.x * b.w) + a.y * b.z) - a.z * b.y, ((a.w * b.y + a.y * b.w) + a.z * b.x) - a.x * b.z,
            ((a.w * b.z + a.z * b.w) + a.x * b.y) - a.y * b.x, ((a.w * b.w - a.x * b.x) - a.y * b.y) - a.z * b.z};
}
// detail::exp (integrate.h:21-32)
__device__ __forceinline__ dq exp_rotation(d3 ang_vel, double dt)
{
    const d3 angle = ang_vel * dt;
    const double mag = sqrt(sqnorm(angle));
    if (mag < 1e-12)
    {
        const d3 half = angle * 0.5;
        const double n = sqrt(((half.x * half.x + half.y * half.y) + half.z * half.z) + 1.0);
        return {half.x / n, half.y / n, half.z / n, 1.0 / n};
    }
    const d3 axis{angle.x / mag, angle.y / mag, angle.z / mag};
    const double s = sin(0.5 * mag), c = cos(0.5 * mag); // Eigen AngleAxis → Quaternion (lin_alg.h:569-576)
    return {s * axis.x, s * axis.y, s * axis.z, c};
}

struct DynArrays
{
    double *vel;      // [n][3]
    double *ang_vel;  // [n][3]
    double *acc;      // [n][3]  M_acc: accumulated force · inv_mass
    double *torque;   // [n][3]  M_torque_acc
    double *mass;     // [n][2]  mass, inv_mass
    double *inertia;  // [n][18] local tensor, local inverse tensor (row-major)
    double *inertia_w; // [n][18] world tensor, world inverse tensor: update_derived_state
};

__device__ __forceinline__ dm3 load_m3(const double *p)
{
    dm3 a;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) a.m[i][j] = p[3 * i + j];
    return a;
}
__device__ __forceinline__ void store_m3(double *p, const dm3 &a)
{
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) p[3 * i + j] = a.m[i][j];
}
// particle::update_derived_state (particle.h:140-146)
__device__ __forceinline__ void derive_state(const double *__restrict__ quat, const DynArrays &dy, uint32_t i)
{
    const dq q{quat[4ull * i], quat[4ull * i + 1], quat[4ull * i + 2], quat[4ull * i + 3]};
    const dm3 r = to_rotation_matrix(q), rt = transpose(r);
    store_m3(dy.inertia_w + 18ull * i, mul(mul(r, load_m3(dy.inertia + 18ull * i)), rt));
    store_m3(dy.inertia_w + 18ull * i + 9, mul(mul(r, load_m3(dy.inertia + 18ull * i + 9)), rt));
}

__global__ void __launch_bounds__(256)
dynamics_derive_kernel(const double *__restrict__ quat, DynArrays dy, uint32_t first, uint32_t count)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) derive_state(quat, dy, first + k);
}

// loop A of world::step_impl (src/world.cpp:22-34) without the broad-phase call, which pk_collide makes:
// apply_force(gravity · mass), integrate_vel, disp = vel · dt, clear_forces.
__global__ void __launch_bounds__(256)
integrate_vel_kernel(const uint8_t *__restrict__ flags, DynArrays dy, double *__restrict__ disp, uint32_t n, double dt, d3 gravity)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t fl = flags[i];
    if (!(fl & FLAG_ALIVE) || (fl & FLAG_STATIC)) return; // slot.available() || is_static()
    const double mass = dy.mass[2ull * i], inv_mass = dy.mass[2ull * i + 1];
    d3 acc{dy.acc[3ull * i], dy.acc[3ull * i + 1], dy.acc[3ull * i + 2]};
    acc = acc + (gravity * mass) * inv_mass; // particle.h:78-79
    d3 vel{dy.vel[3ull * i], dy.vel[3ull * i + 1], dy.vel[3ull * i + 2]};
    vel = vel + acc * dt; // integrate.h:38
    d3 w{dy.ang_vel[3ull * i], dy.ang_vel[3ull * i + 1], dy.ang_vel[3ull * i + 2]};
    const d3 torque{dy.torque[3ull * i], dy.torque[3ull * i + 1], dy.torque[3ull * i + 2]};
    const dm3 iw = load_m3(dy.inertia_w + 18ull * i), iiw = load_m3(dy.inertia_w + 18ull * i + 9);
    const d3 alpha = mul(iiw, torque - cross(w, mul(iw, w))); // angular_accel, particle.h:72-76
    w = w + alpha * dt; // integrate.h:39
    const d3 dsp = vel * dt; // src/world.cpp:30-31
    dy.vel[3ull * i] = vel.x; dy.vel[3ull * i + 1] = vel.y; dy.vel[3ull * i + 2] = vel.z;
    dy.ang_vel[3ull * i] = w.x; dy.ang_vel[3ull * i + 1] = w.y; dy.ang_vel[3ull * i + 2] = w.z;
    disp[3ull * i] = dsp.x; disp[3ull * i + 1] = dsp.y; disp[3ull * i + 2] = dsp.z;
    dy.acc[3ull * i] = 0.0; dy.acc[3ull * i + 1] = 0.0; dy.acc[3ull * i + 2] = 0.0; // clear_forces, particle.h:101-105
    dy.torque[3ull * i] = 0.0; dy.torque[3ull * i + 1] = 0.0; dy.torque[3ull * i + 2] = 0.0;
}

// loop B (src/world.cpp:50-55): integrate_pos (integrate.h:42-46); the orientation setter refreshes the
// derived state (particle.h:40-44).
__global__ void __launch_bounds__(256)
integrate_pos_kernel(const uint8_t *__restrict__ flags, DynArrays dy, double *__restrict__ pos, double *__restrict__ quat, uint32_t n, double dt)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t fl = flags[i];
    if (!(fl & FLAG_ALIVE) || (fl & FLAG_STATIC)) return;
    const d3 vel{dy.vel[3ull * i], dy.vel[3ull * i + 1], dy.vel[3ull * i + 2]};
    const d3 w{dy.ang_vel[3ull * i], dy.ang_vel[3ull * i + 1], dy.ang_vel[3ull * i + 2]};
    d3 p{pos[3ull * i], pos[3ull * i + 1], pos[3ull * i + 2]};
    p = p + vel * dt;
    const dq q{quat[4ull * i], quat[4ull * i + 1], quat[4ull * i + 2], quat[4ull * i + 3]};
    const dq qn = qmul(exp_rotation(w, dt), q);
    pos[3ull * i] = p.x; pos[3ull * i + 1] = p.y; pos[3ull * i + 2] = p.z;
    quat[4ull * i] = qn.x; quat[4ull * i + 1] = qn.y; quat[4ull * i + 2] = qn.z; quat[4ull * i + 3] = qn.w;
    derive_state(quat, dy, i);
}

} // namespace pk
