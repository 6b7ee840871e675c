"""The oracle's integrator restatement (detail/integrate.h:17-47, core/particle.h:72-105, 140-146,
src/world.cpp:22-34, 50-55).  The reference has no test of its own for this code, so it is checked against
the closed forms the code implements ("parity unpinned" at the ulp level)."""
import numpy as np

import oracle

DT = 1.0 / 60.0


def _one(pos=(0, 0, 0), quat=(0, 0, 0, 1), vel=(0, 0, 0), w=(0, 0, 0), mass=1.0, inertia=None, flags=2):
    inertia = np.eye(3) if inertia is None else np.asarray(inertia, float)
    return oracle.Dynamics([pos], [quat], [vel], [w], [mass], [inertia.ravel()], [flags])


def test_free_fall_is_semi_implicit_euler():  # integrate.h:36-46: v += a·dt, then x += v·dt
    d = _one(pos=(0, 10, 0), vel=(1, 0, 0), mass=3.0)
    g = (0.0, -9.81, 0.0)
    y, v = 10.0, 0.0
    for _ in range(60):
        disp = d.integrate_velocities(DT, g)
        v = v + ((-9.81 * 3.0) * (1.0 / 3.0)) * DT
        assert d.vel[0, 1] == v and disp[0, 1] == v * DT
        d.integrate_positions(DT)
        y = y + v * DT
        assert d.pos[0, 1] == y
    assert abs(d.pos[0, 0] - 1.0) < 1e-12
    assert np.all(d.acc == 0) and np.all(d.torque == 0)  # clear_forces


def test_static_and_dead_bodies_are_skipped():  # src/world.cpp:24, 52
    for fl in (3, 0):
        d = _one(pos=(0, 1, 0), vel=(1, 2, 3), flags=fl)
        d.integrate_velocities(DT, (0, -9.81, 0))
        d.integrate_positions(DT)
        assert np.array_equal(d.pos, [[0, 1, 0]]) and np.array_equal(d.vel, [[1, 2, 3]])


def test_constant_spin_about_z():  # detail::exp → from_angle_axis, premultiplied (integrate.h:21-32, 45)
    w = 2.5
    d = _one(w=(0, 0, w))
    for _ in range(90):
        d.integrate_velocities(DT, (0, 0, 0))
        d.integrate_positions(DT)
    th = w * 90 * DT
    assert np.allclose(d.quat[0], [0, 0, np.sin(th / 2), np.cos(th / 2)], atol=1e-13)
    assert d.ang_vel[0, 2] == w  # symmetric top about its axis: ω × Iω = 0 exactly


def test_small_angle_branch_is_normalised():  # integrate.h:25-30
    d = _one(w=(3e-11, -4e-11, 0))
    d.integrate_positions(DT)
    q = d.quat[0]
    assert abs(np.linalg.norm(q) - 1.0) < 1e-15 and q[3] > 0.999999
    assert np.allclose(q[:3], [0.5 * 3e-11 * DT, -0.5 * 4e-11 * DT, 0], rtol=1e-12, atol=0)


def test_angular_acceleration_is_eulers_equation_in_the_world_frame():  # particle.h:72-76, 140-146
    rng = np.random.default_rng(5)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    x, y, z, w_ = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w_), 2 * (x * z + y * w_)],
                  [2 * (x * y + z * w_), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w_)],
                  [2 * (x * z - y * w_), 2 * (y * z + x * w_), 1 - 2 * (x * x + y * y)]])
    I = np.diag([0.4, 1.1, 2.3])
    om = np.array([0.7, -1.2, 0.5])
    tq = np.array([0.3, 0.1, -0.4])
    d = _one(quat=q, w=om, inertia=I)
    d.torque[:] = tq
    d.integrate_velocities(DT, (0, 0, 0))
    Iw = R @ I @ R.T
    alpha = np.linalg.inv(Iw) @ (tq - np.cross(om, Iw @ om))
    assert np.allclose(d.ang_vel[0], om + alpha * DT, rtol=1e-13, atol=1e-15)


def test_torque_free_tumbling_conserves_angular_momentum_to_first_order():
    I = np.diag([0.5, 1.0, 2.0])
    d = _one(w=(1.0, 0.2, 0.1), inertia=I)
    L0 = I @ d.ang_vel[0]
    dt = 1e-4
    for _ in range(2000):
        d.integrate_velocities(dt, (0, 0, 0))
        d.integrate_positions(dt)
    x, y, z, w_ = d.quat[0]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w_), 2 * (x * z + y * w_)],
                  [2 * (x * y + z * w_), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w_)],
                  [2 * (x * z - y * w_), 2 * (y * z + x * w_), 1 - 2 * (x * x + y * y)]])
    L = R @ I @ R.T @ d.ang_vel[0]
    assert np.linalg.norm(L - L0) < 2e-3 * np.linalg.norm(L0)  # explicit first-order scheme: O(dt) drift
