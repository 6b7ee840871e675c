"""-m gpu parity of the device integrator (detail/integrate.h:17-47 and the per-body loops of
src/world.cpp:22-34, 50-55) against the oracle's restatement.

Tolerances, stated: loop A (velocities, displacement) is mul / add only and must be bit-identical.  Loop B's
positions are bit-identical; the orientation goes through sin / cos (libdevice vs libm: up to 2 ulp apart)
and is compared to 1e-14 absolute per step, 1e-11 after a 120-step replay (the world tensors inherit it)."""
import numpy as np
import pytest

import oracle
import physkit_b200 as pk
from scenes import SplitMix64, scene_c1, scene_c3

pytestmark = pytest.mark.gpu

G = (0.0, -9.81, 0.0)
DT = 1.0 / 60.0


def _state(sc, seed):
    rng = SplitMix64(seed)
    n = sc.n
    vel = rng.uniform(-2, 2, n, 3)
    w = rng.uniform(-3, 3, n, 3)
    w[::5] = 0.0  # exp()'s small-angle branch (integrate.h:25-30)
    w[1::5] *= 1e-14
    mass = rng.uniform(0.5, 4.0, n)
    a = rng.uniform(-0.2, 0.2, n, 3, 3)
    inertia = np.einsum("nij,nkj->nik", a, a) + np.eye(3) * rng.uniform(0.3, 2.0, n)[:, None, None]  # SPD, full
    return vel, w, mass, inertia.reshape(n, 9)


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def test_velocity_loop_bit_exact_and_feeds_the_fat_boxes():
    from gpu_util import make_context

    sc = scene_c1(side=6)
    vel, w, mass, inertia = _state(sc, 1)
    ctx = make_context(sc, max_pairs=200_000, mode=pk.MODE_WORLD)
    ctx.dynamics_enable()
    ctx.dynamics_upload(vel, w, mass, inertia)
    ref = oracle.Dynamics(sc.pos, sc.quat, vel, w, mass, inertia, sc.flags)
    world = oracle.World(sc.shapes)
    rng = SplitMix64(2)
    for step in range(3):
        acc = rng.uniform(-1, 1, sc.n, 3)
        tq = rng.uniform(-1, 1, sc.n, 3)
        ref.acc[:], ref.torque[:] = acc, tq
        ctx.dynamics_set_forces(acc, tq)
        disp = ref.integrate_velocities(DT, G)
        ctx.integrate_velocities(DT, G)
        pos, quat, v, om = ctx.dynamics_download(sc.n)
        dyn = (sc.flags & 1) == 0
        assert np.array_equal(_bits(v), _bits(ref.vel))
        if step == 0:  # the world tensors are bit-identical as long as the orientation is (no sin / cos yet)
            assert np.array_equal(_bits(om), _bits(ref.ang_vel))
        else:
            assert np.allclose(om, ref.ang_vel, rtol=1e-12, atol=1e-13)
        assert np.array_equal(_bits(ctx.displacements(sc.n)[dyn]), _bits(disp[dyn]))
        assert np.array_equal(v[~dyn], vel[~dyn])  # static bodies are skipped (src/world.cpp:24)
        # the displacement reaches broad_phase::update_node: same fat boxes, same pair set
        world.step(ref.pos, ref.quat, disp, sc.shape_id, sc.flags)
        ctx.collide()
        assert np.array_equal(ctx.pairs(), world.pairs())
        ref.integrate_positions(DT)
        ctx.integrate_positions(DT)
        pos, quat, v, om = ctx.dynamics_download(sc.n)
        assert np.array_equal(_bits(pos), _bits(ref.pos))
        assert np.abs(quat - ref.quat).max() < 1e-14
        ref.quat[:] = quat  # keep both sides on the same orientation so the next step's linear part stays exact
    ctx.close()


def test_free_flight_replay_stays_within_tolerance():
    from gpu_util import make_context

    sc = scene_c3(side=16)
    vel, w, mass, inertia = _state(sc, 7)
    ctx = make_context(sc, max_pairs=600_000, mode=pk.MODE_WORLD)
    ctx.dynamics_enable()
    ctx.dynamics_upload(vel, w, mass, inertia)
    ref = oracle.Dynamics(sc.pos, sc.quat, vel, w, mass, inertia, sc.flags)
    for _ in range(120):
        ref.integrate_velocities(DT, G)
        ref.integrate_positions(DT)
        ctx.integrate_velocities(DT, G)
        ctx.integrate_positions(DT)
    pos, quat, v, om = ctx.dynamics_download(sc.n)
    assert np.array_equal(_bits(v), _bits(ref.vel))  # gravity only: the linear part never sees the orientation
    assert np.array_equal(_bits(pos), _bits(ref.pos))
    assert np.abs(quat - ref.quat).max() < 1e-11
    assert np.allclose(om, ref.ang_vel, rtol=1e-10, atol=1e-11)
    assert np.abs(np.linalg.norm(quat, axis=1) - 1.0).max() < 1e-9  # exp() is a unit quaternion; no renormalisation in the reference either
    ctx.close()


def test_call_order_and_infinite_mass():
    ctx = pk.Context(4, 64, mode=pk.MODE_WORLD, max_shapes=2)
    sid = ctx.add_shapes([("obb", np.array([0.5, 0.5, 0.5]))])[0]
    with pytest.raises(pk.PkError) as e:
        ctx.integrate_velocities(DT, G)
    assert e.value.status == -7
    ctx.resize(2)
    pos = np.array([[0.0, 5, 0], [3.0, 5, 0]])
    quat = np.tile([0.0, 0, 0, 1], (2, 1))
    ctx.upload(pos, quat, np.zeros((2, 3)), np.array([sid, sid], np.uint32), np.array([2, 2], np.uint8))
    ctx.dynamics_enable()
    ctx.dynamics_upload(np.zeros((2, 3)), np.array([[0, 0, 1.0], [0, 0, 1.0]]), np.array([2.0, np.inf]), np.tile(np.eye(3).ravel(), (2, 1)))
    ctx.dynamics_set_forces(np.zeros((2, 3)), np.array([[0, 0, 4.0], [0, 0, 4.0]]))
    ctx.integrate_velocities(DT, G)
    _, _, v, om = ctx.dynamics_download(2)
    assert v[0, 1] == (0.0 + (-9.81 * 2.0) * 0.5) * DT
    assert om[0, 2] == 1.0 + 4.0 * DT
    # infinite mass: inv_mass = 0, zero inverse tensor (particle.h:25-26) → gravity·inf·0 = NaN in the reference too
    assert om[1, 2] == 1.0
    ctx.close()
