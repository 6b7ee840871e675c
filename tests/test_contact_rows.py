"""Contact Jacobian set-up: constraint_solver::setup_contacts (reference
include/physkit/collision/constraint.h:104-113, 874-953, 1052-1104).

CPU: the oracle's restatement against the closed forms the code implements (the reference has no test of its
own at this level).  GPU: pk_contact_rows_setup against the restatement over a moving-pile replay, bit for bit
(no transcendental function on this path: mul / add / div / sqrt in the reference's order, no FMA)."""
import numpy as np
import pytest

import oracle
from scenes import SplitMix64, scene_c3
from test_manifolds import _fake_solver_impulses, _oracle_step, _replay

DT = 1.0 / 60.0


def _body_state(n, seed):
    rng = SplitMix64(seed)
    vel = rng.uniform(-1.5, 1.5, n, 3)
    w = rng.uniform(-2, 2, n, 3)
    mass = rng.uniform(0.5, 4.0, n)
    a = rng.uniform(-0.2, 0.2, n, 3, 3)
    inertia = (np.einsum("nij,nkj->nik", a, a) + np.eye(3) * rng.uniform(0.3, 2.0, n)[:, None, None]).reshape(n, 9)
    rest = rng.uniform(0.0, 1.0, n)
    fric = rng.uniform(0.1, 1.0, n)
    return vel, w, mass, inertia, rest, fric


def _two_body_manifold(depth=0.02, normal=(0.0, 1.0, 0.0)):
    """Body 1 rests on body 2 with one contact point; returns (Manifolds, pos, quat)."""
    M = oracle.Manifolds()
    pos = np.array([[0.0, 0, 0], [0.0, 1.0, 0.0], [0.0, 0.0, 0.0]])
    quat = np.tile([0.0, 0.0, 0.0, 1.0], (3, 1))
    key = np.array([(1 << 32) | 2], dtype=np.uint64)
    n = np.asarray(normal, float)
    wa = np.array([0.1, 0.5, 0.05])           # on body a (id 1)
    wb = wa + n * depth                         # on body b (id 2): b.pos + r_b − (a.pos + r_a) = n·depth
    c = np.concatenate([n, wa, wb, [depth]])[None, :]
    M.step(key, np.array([1], np.uint8), c, pos, quat)
    return M, pos, quat


def test_rows_are_the_contact_jacobian():
    M, pos, quat = _two_body_manifold()
    vel = np.array([[0.0, 0, 0], [0.0, -3.0, 0.0], [0.0, 0.0, 0.0]])
    w = np.zeros((3, 3))
    mass = np.array([1.0, 2.0, 4.0])
    inertia = np.tile(np.eye(3).ravel(), (3, 1)) * np.array([1.0, 0.5, 2.0])[:, None]
    rest = np.array([0.0, 0.3, 0.6])
    fric = np.array([0.5, 0.4, 0.9])
    keys, pts, rows = M.setup_contacts(pos, quat, vel, w, mass, inertia, rest, fric, DT, 9.81)
    assert len(keys) == 1 and pts[0] == 0
    r = rows[0]
    n = r[0:3]
    assert np.allclose(n, [0, 1, 0])
    r_a = np.array([0.1, 0.5, 0.05]) - pos[1]
    r_b = np.array([0.1, 0.52, 0.05]) - pos[2]
    assert np.allclose(r[3:6], np.cross(r_a, n)) and np.allclose(r[6:9], -np.cross(r_b, n))
    k = 1 / 2.0 + 1 / 4.0 + r[3:6] @ (r[3:6] / 0.5) + r[6:9] @ (r[6:9] / 2.0)
    assert np.isclose(r[9], 1.0 / k, rtol=1e-14)
    # approaching at 3 m/s > threshold 2·g·dt: restitution bias = max(e_a, e_b)·v_rel_n wins over Baumgarte
    v_rel = (vel[1] - vel[2]) @ n
    assert np.isclose(r[10], min((-0.1 / DT) * (0.02 - 0.005), 0.6 * v_rel))
    # tangents: orthonormal to n (build_orthonormal_basis), no bias
    t1, t2 = r[11:14], r[22:25]
    assert abs(t1 @ n) < 1e-15 and abs(t2 @ n) < 1e-15 and abs(t1 @ t2) < 1e-15
    assert np.isclose(np.linalg.norm(t1), 1) and np.isclose(np.linalg.norm(t2), 1)
    assert r[21] == 0.0 and r[32] == 0.0
    assert np.isclose(r[33], np.sqrt(0.4 * 0.9))
    # the 2×2 friction block inverse
    m11, m22 = 1 / r[20], 1 / r[31]
    m12 = r[14:17] @ (r[25:28] / 0.5) + r[17:20] @ (r[28:31] / 2.0)
    inv = np.linalg.inv(np.array([[m11, m12], [m12, m22]]))
    assert np.allclose([r[34], r[35], r[36]], [inv[0, 0], inv[0, 1], inv[1, 1]], rtol=1e-12)


def test_slow_contact_uses_baumgarte_only_and_separated_points_yield_no_row():
    M, pos, quat = _two_body_manifold(depth=0.03)
    z = np.zeros((3, 3))
    one = np.ones(3)
    eye = np.tile(np.eye(3).ravel(), (3, 1))
    keys, pts, rows = M.setup_contacts(pos, quat, z, z, one, eye, one * 0.5, one * 0.5, DT, 9.81)
    assert np.isclose(rows[0, 10], (-0.1 / DT) * max(0.0, 0.03 - 0.005), rtol=1e-12)  # resting: v_rel_n = 0 ≥ −threshold
    # move b away along the normal: penetration <= 0 → build_contact_jacobian returns nullopt (:893-894)
    pos2 = pos.copy()
    pos2[2, 1] -= 0.05
    keys, pts, rows = M.setup_contacts(pos2, quat, z, z, one, eye, one * 0.5, one * 0.5, DT, 9.81)
    assert len(keys) == 0


def test_orthonormal_basis_is_branch_free_at_the_poles():  # constraint.h:104-113
    for nz in (1.0, -1.0):
        M, pos, quat = _two_body_manifold(normal=(0.0, 0.0, nz))
        z = np.zeros((3, 3))
        one = np.ones(3)
        eye = np.tile(np.eye(3).ravel(), (3, 1))
        _, _, rows = M.setup_contacts(pos, quat, z, z, one, eye, one * 0.5, one * 0.5, DT, 9.81)
        t1, t2, n = rows[0, 11:14], rows[0, 22:25], rows[0, 0:3]
        assert np.all(np.isfinite(rows[0])) and abs(t1 @ t2) < 1e-15 and abs(t1 @ n) < 1e-15 and abs(t2 @ n) < 1e-15


@pytest.mark.gpu
def test_gpu_contact_rows_match_the_oracle_over_a_replay():
    import physkit_b200 as pk
    from gpu_util import make_context

    state = {}
    M = oracle.Manifolds()
    totals = {"rows": 0, "skipped": 0}

    def step(k, sc, pos, disp):
        if "ctx" not in state:
            state["w"] = oracle.World(sc.shapes)
            ctx = state["ctx"] = make_context(sc, max_pairs=200_000, mode=pk.MODE_WORLD)
            ctx.manifolds_enable(20_000)
            ctx.dynamics_enable()
            state["dyn"] = _body_state(sc.n, 99)
            vel, w, mass, inertia, rest, fric = state["dyn"]
            ctx.dynamics_upload(vel, w, mass, inertia)
            ctx.material_upload(rest, fric)
        ctx = state["ctx"]
        vel, w, mass, inertia, rest, fric = state["dyn"]
        keys, hit, began, ended = _oracle_step(state["w"], M, sc, pos, disp)
        ctx.upload(pos, sc.quat, disp, sc.shape_id, sc.flags)
        ctx.collide()
        ctx.manifolds_update()
        # a host solver's velocities of this step (here: a deterministic perturbation)
        vel = vel + 0.05 * np.sin(np.arange(sc.n)[:, None] * 0.37 + k)
        ctx.dynamics_set_velocities(vel, w)
        n = ctx.contact_rows_setup(DT, 9.81)
        got = ctx.contact_rows()
        rk, rp, rows = M.setup_contacts(pos, sc.quat, vel, w, mass, inertia, rest, fric, DT, 9.81)
        mk, mc, mp = M.get()
        assert n == len(rk) == len(got), (n, len(rk))
        assert np.array_equal(got["key"], rk) and np.array_equal(got["point"], rp)
        assert np.array_equal(mk[got["manifold"]], rk)
        flat = np.concatenate([np.concatenate([got[r]["J_v"], got[r]["J_w_a"], got[r]["J_w_b"], got[r]["M_eff"][:, None], got[r]["bias"][:, None]], axis=1)
                               for r in ("normal", "tangent1", "tangent2")] +
                              [got["friction_coeff"][:, None], got["inv_m_11"][:, None], got["inv_m_12"][:, None], got["inv_m_22"][:, None],
                               got["accumulated"]], axis=1)
        assert np.array_equal(np.ascontiguousarray(flat).view(np.uint64), np.ascontiguousarray(rows).view(np.uint64)), f"step {k}"
        totals["rows"] += n
        totals["skipped"] += int(mc.sum()) - n
        imp = _fake_solver_impulses(mk, mc, k)
        M.set_impulses(imp)
        ctx.manifolds_set_impulses(imp)
        if k > 1:  # step 0 has no pairs (first-step quirk), step 1 starts the manifolds, step 2 sees their impulses
            assert np.any(got["accumulated"] != 0.0)  # warm-start impulses travel into the rows

    try:
        _replay(10, 9, step)
    finally:
        if "ctx" in state:
            state["ctx"].close()
    assert totals["rows"] > 2000 and totals["skipped"] > 0


@pytest.mark.gpu
def test_gpu_contact_rows_call_order():
    import physkit_b200 as pk
    from gpu_util import make_context

    sc = scene_c3(side=4)
    ctx = make_context(sc, max_pairs=10_000, mode=pk.MODE_WORLD)
    try:
        with pytest.raises(pk.PkError) as e:
            ctx.contact_rows_setup(DT)
        assert e.value.status == -7
        ctx.manifolds_enable(1000)
        ctx.dynamics_enable()
        ctx.collide()
        ctx.manifolds_update()
        assert ctx.contact_rows_setup(DT) == 0 and len(ctx.contact_rows()) == 0  # first step of a world: no pairs yet
    finally:
        ctx.close()
