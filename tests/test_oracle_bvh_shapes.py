"""Pin the CPU oracle's broadphase / shape layer against the reference's own tests:
tests/dynamic_bvh/main.cpp, tests/obb/obb_test.cpp, tests/mesh/main.cpp (cited per test), plus the
equivalence proof the GPU design rests on: the persistent pair set maintained by
broad_phase::calculate_pairs equals a stateless function of the stored boxes and move history."""
import math

import numpy as np

import oracle
from scenes import IDENT, SplitMix64, angle_axis, scene_c1, scene_c2


def make_box(cx, cy, cz, half=0.5):
    return np.array([cx - half, cy - half, cz - half, cx + half, cy + half, cz + half], dtype=np.float64)


def intersects(a, b):
    return bool(np.all(a[:3] <= b[3:]) and np.all(a[3:] >= b[:3]))


# ------------------------------------------------------------------ dynamic_bvh (main.cpp)
def test_add_single_leaf():  # main.cpp:24-32
    t = oracle.DynamicBVH()
    box = make_box(0, 0, 0)
    h = t.add(0, box)
    assert t.data(h) == 0
    assert np.array_equal(t.bounds(h), box)  # add() stores the exact box (bvh.h:294-302)


def test_add_returns_distinct_handles_and_preserves_data():  # main.cpp:34-58
    t = oracle.DynamicBVH()
    hs = [t.add(i, make_box(i * 2.0, 0, 0)) for i in range(20)]
    assert len(set(hs)) == 20
    assert [t.data(h) for h in hs] == list(range(20))
    assert t.validate()


def test_remove_variants():  # main.cpp:64-146
    t = oracle.DynamicBVH()
    h = t.add(0, make_box(0, 0, 0))
    t.remove_leaf(h)
    assert len(t.query_aabb(make_box(0, 0, 0, 100))) == 0
    t = oracle.DynamicBVH()
    h0 = t.add(0, make_box(0, 0, 0))
    t.add(1, make_box(5, 0, 0))
    t.remove_leaf(h0)
    assert set(t.query_aabb(make_box(0, 0, 0, 100))) == {1}
    t = oracle.DynamicBVH()
    hs = [t.add(i, make_box(i * 2.0, 0, 0)) for i in range(10)]
    for i in range(0, 10, 2):
        t.remove_leaf(hs[i])
    assert set(t.query_aabb(make_box(0, 0, 0, 100))) == {1, 3, 5, 7, 9}
    assert t.validate()
    for i in range(1, 10, 2):
        t.remove_leaf(hs[i])
    assert len(t.query_aabb(make_box(0, 0, 0, 1000))) == 0


def test_update_within_fat_bounds_no_reinsert():  # main.cpp:149-164 (margin 0.1)
    t = oracle.DynamicBVH()
    h = t.add(0, make_box(0, 0, 0))
    assert t.update_leaf(h, make_box(0.01, 0, 0), (0.01, 0, 0))
    assert not t.update_leaf(h, make_box(0.02, 0, 0), (0, 0, 0))


def test_update_fat_box_values():  # bvh.cpp:487-506, exact arithmetic of the fat rule
    t = oracle.DynamicBVH()
    h = t.add(0, make_box(0, 0, 0))
    tb = make_box(5, 0, 0)
    assert t.update_leaf(h, tb, (2.0, -0.25, 0.0))
    fat = t.bounds(h)
    exp = np.array([tb[0] - 0.1, (tb[1] - 0.1) + -0.25, tb[2] - 0.1, (tb[3] + 0.1) + 2.0, tb[4] + 0.1, (tb[5] + 0.1) + 0.0])
    assert np.array_equal(fat, exp)
    assert fat[3] > tb[3]  # main.cpp:196-208 predictive expansion


def test_update_outside_fat_bounds_reinserts_and_preserves_data():  # main.cpp:166-194
    t = oracle.DynamicBVH()
    h = t.add(42, make_box(0, 0, 0))
    assert t.update_leaf(h, make_box(10, 10, 10), (10, 10, 10))
    assert 42 in set(t.query_aabb(make_box(10, 10, 10, 2)))
    assert t.data(h) == 42


def test_query_basic():  # main.cpp:214-277
    t = oracle.DynamicBVH()
    assert len(t.query_aabb(make_box(0, 0, 0, 100))) == 0
    t.add(0, make_box(0, 0, 0))
    t.add(1, make_box(10, 10, 10))
    t.add(2, make_box(0.5, 0.5, 0.5))
    f = set(t.query_aabb(make_box(0.25, 0.25, 0.25, 1)))
    assert f == {0, 2}
    assert len(t.query_aabb(make_box(100, 100, 100, 0.5))) == 0
    t = oracle.DynamicBVH()
    for i in range(10):
        t.add(i, make_box(0, 0, 0))
    assert len(t.query_aabb(make_box(0, 0, 0, 2), stop_after=1)) == 1  # early termination


def _crosscheck(n, seed, pos_rng, size_rng, nq):
    rng = SplitMix64(seed)
    t = oracle.DynamicBVH()
    boxes = []
    for i in range(n):
        half = rng.uniform(*size_rng)
        b = make_box(rng.uniform(*pos_rng), rng.uniform(*pos_rng), rng.uniform(*pos_rng), half)
        t.add(i, b)
        boxes.append(b)
    assert t.validate()
    for _ in range(nq):
        q = make_box(rng.uniform(*pos_rng), rng.uniform(*pos_rng), rng.uniform(*pos_rng), rng.uniform(*size_rng) * 5)
        got = set(t.query_aabb(q))
        want = {i for i in range(n) if intersects(boxes[i], q)}
        assert got == want  # exact boxes ⇒ not just a superset


def test_query_large_population():  # main.cpp:279-311 (n=200)
    _crosscheck(200, 12345, (-50, 50), (0.5, 0.5), 4)


def test_query_aabb_brute_force_crosscheck():  # main.cpp:600-635 (n=100, 20 queries)
    _crosscheck(100, 2025, (-30, 30), (0.2, 2.0), 20)


def test_heavy_churn():  # main.cpp:483-522
    rng = SplitMix64(99)
    t = oracle.DynamicBVH()
    hs = [t.add(i, make_box(*rng.uniform(-20, 20, 3))) for i in range(50)]
    order = np.argsort(rng.u01(50))
    removed = set()
    for k in order[:25]:
        removed.add(t.data(hs[k]))
        t.remove_leaf(hs[k])
    for i in range(50, 75):
        t.add(i, make_box(*rng.uniform(-20, 20, 3)))
    found = set(t.query_aabb(make_box(0, 0, 0, 1000)))
    assert len(found) == 50 and not (found & removed)
    assert t.validate()


def test_edge_cases():  # main.cpp:687-718
    t = oracle.DynamicBVH()
    t.add(0, np.array([0, 0, 0, 1, 1, 0], dtype=np.float64))
    assert 0 in set(t.query_aabb(make_box(0.5, 0.5, 0, 2)))
    t = oracle.DynamicBVH()
    box = make_box(0, 0, 0)
    for i in range(5):
        t.add(i, box)
    assert len(set(t.query_aabb(box))) == 5


def test_interleaved_operations():  # main.cpp:787-856 (500 random ops)
    rng = SplitMix64(1337)
    t = oracle.DynamicBVH()
    handles, live, next_id = [], set(), 0
    for _ in range(500):
        op = 0 if not handles else int(rng.randint(1, 4)[0])
        if op == 0:
            h = t.add(next_id, make_box(*rng.uniform(-30, 30, 3)))
            handles.append(h)
            live.add(next_id)
            next_id += 1
        elif op == 1:
            k = int(rng.randint(1, len(handles))[0])
            live.discard(t.data(handles[k]))
            t.remove_leaf(handles[k])
            handles.pop(k)
        elif op == 2:
            k = int(rng.randint(1, len(handles))[0])
            t.update_leaf(handles[k], make_box(*rng.uniform(-30, 30, 3)), rng.uniform(-30, 30, 3))
        else:
            assert live <= set(t.query_aabb(make_box(0, 0, 0, 1000)))
        assert t.validate()
    assert set(t.query_aabb(make_box(0, 0, 0, 1000))) == live


# ------------------------------------------------------------------ shapes (obb_test / mesh main)
def test_aabb_intersects_inclusive():  # tests/mesh/main.cpp:109-121
    a = np.array([0, 0, 0, 1, 1, 1.0])
    b = np.array([0.5, 0.5, 0.5, 1.5, 1.5, 1.5])
    c = np.array([2, 2, 2, 3, 3, 3.0])
    d = np.array([1, 1, 1, 2, 2, 2.0])  # corner touch counts
    for x, y, want in [(a, b, 1), (a, c, 0), (a, d, 1)]:
        assert len(oracle.brute_pairs(np.stack([x, y]))) == want
        assert len(oracle.query_pairs(np.stack([x, y]))) == want


def test_obb_support_point():  # tests/obb/obb_test.cpp:142-159
    s = oracle.support([("obb", (1, 2, 3))], (0, 0, 0), IDENT, 0, (1.0, -2.0, 0.3))
    assert np.allclose(s, (1.0, -2.0, 3.0), atol=1e-9)
    rot = angle_axis(math.pi / 2.0, (0, 0, 1))
    s = oracle.support([("obb", (2, 1, 1))], (1, 1, 0), rot, 0, (1.0, 0.0, 0.0))
    assert np.allclose(s, (2.0, 3.0, 1.0), atol=1e-9)


CUBE01 = np.array(
    [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=np.float64
)  # tests/mesh/main.cpp:15-26 cube_fixture


def test_aabb_quat_transform():  # tests/mesh/main.cpp:179-201
    rot = angle_axis(math.pi / 2.0, (0, 0, 1))
    b = oracle.bounds([("hull", CUBE01)], [(0, 0, 0)], [rot], [0])[0]
    assert np.allclose(b, (-1, 0, 0, 0, 1, 1), atol=1e-9)


def test_mesh_support():  # tests/mesh/main.cpp:890-899
    sh = [("hull", CUBE01)]
    assert abs(oracle.support(sh, (0, 0, 0), IDENT, 0, (1, 0, 0))[0] - 1.0) < 1e-9
    assert abs(oracle.support(sh, (0, 0, 0), IDENT, 0, (-1, 0, 0))[0] - 0.0) < 1e-9
    assert abs(oracle.support(sh, (0, 0, 0), IDENT, 0, (0, 1, 0))[1] - 1.0) < 1e-9
    assert np.allclose(oracle.support(sh, (0, 0, 0), IDENT, 0, (1, 1, 1)), (1, 1, 1), atol=1e-9)
    # strict '>' ⇒ lowest index wins ties (src/mesh.cpp:350): d=(1,0,0) ties verts 1,2,5,6 → 1
    assert np.array_equal(oracle.support(sh, (0, 0, 0), IDENT, 0, (1, 0, 0)), CUBE01[1])


def test_mesh_instance_bounds_support_rotated():  # tests/mesh/main.cpp:1620-1710
    sh = [("hull", CUBE01)]
    b = oracle.bounds(sh, [(5, 0, 0)], [IDENT], [0])[0]
    assert np.allclose(b, (5, 0, 0, 6, 1, 1), atol=1e-9)
    assert abs(oracle.support(sh, (5, 0, 0), IDENT, 0, (1, 0, 0))[0] - 6.0) < 1e-9
    rot = angle_axis(math.pi / 2.0, (0, 0, 1))
    b = oracle.bounds(sh, [(0, 0, 0)], [rot], [0])[0]
    assert abs(b[0] + 1.0) < 1e-9 and abs(b[3]) < 1e-9


# ------------------------------------------------------------------ pair-set semantics
def test_query_mode_equals_brute_force():
    sc = scene_c2(3000)
    boxes = oracle.bounds(sc.shapes, sc.pos, sc.quat, sc.shape_id)
    assert np.array_equal(oracle.query_pairs(boxes), oracle.brute_pairs(boxes))


def _stateless_pairs(stored, alive, last_move, create):
    """active(a,b) ⇔ intersects(stored a, stored b) ∧
    (last_move[a] ≥ create[b] ∨ last_move[b] ≥ create[a])."""
    ids = np.nonzero(alive)[0]
    keys = []
    for ii, a in enumerate(ids):
        for b in ids[ii + 1 :]:
            if intersects(stored[a], stored[b]) and (last_move[a] >= create[b] or last_move[b] >= create[a]):
                keys.append((int(a) << 32) | int(b))
    return np.array(sorted(keys), dtype=np.uint64)


def test_world_first_step_quirk_and_stateless_equivalence():
    """Step 1 yields zero pairs (add() stores exact boxes, nothing moves: collision_phases.h:342-346,
    src/world.cpp:30 vs :54); afterwards the faithful incremental pair set must equal the stateless
    rule, including bodies created / destroyed mid-run and static bodies."""
    sc = scene_c1(side=5, spacing=1.02)
    n = sc.n
    rng = SplitMix64(7)
    w = oracle.World(sc.shapes)
    pos = sc.pos.copy()
    flags = sc.flags.copy()
    late = np.zeros(n, dtype=bool)
    late[n // 2 :: 7] = True  # created at step 4
    flags[late] = 0
    stored = np.zeros((n, 6))
    alive = np.zeros(n, dtype=bool)
    last_move = np.full(n, -1, dtype=np.int64)
    create = np.zeros(n, dtype=np.int64)
    vel = np.zeros((n, 3))
    dt = 1.0 / 60.0
    total_pairs = 0
    for step in range(40):
        if step == 4:
            flags[late] = 2
        if step == 9:
            flags[3:40:5] = 0  # destroy a few
        if step == 15:
            flags[3:40:5] = 2  # slots reused
        dyn = (flags & 1) == 0
        vel[dyn, 1] -= 9.81 * dt
        if step % 3 == 0:
            vel[dyn] += rng.uniform(-0.5, 0.5, n, 3)[dyn]
        disp = vel * dt
        moved = w.step(pos, sc.quat, disp, sc.shape_id, flags)
        keys = w.pairs()
        # replay the stateless bookkeeping
        true_box = oracle.bounds(sc.shapes, pos, sc.quat, sc.shape_id)
        n_moved = 0
        for i in range(n):
            now = bool(flags[i] & 2)
            if alive[i] and not now:
                alive[i] = False
                last_move[i] = -1
            if now and not alive[i]:
                alive[i] = True
                stored[i] = true_box[i]
                create[i] = step
                last_move[i] = -1
            if alive[i] and not (flags[i] & 1):
                s, tb = stored[i], true_box[i]
                if not (np.all(tb[:3] >= s[:3]) and np.all(tb[3:] <= s[3:])):
                    nb = np.concatenate([tb[:3] - 0.1, tb[3:] + 0.1])
                    for k in range(3):
                        if disp[i, k] < 0.0:
                            nb[k] = nb[k] + disp[i, k]
                        else:
                            nb[3 + k] = nb[3 + k] + disp[i, k]
                    stored[i] = nb
                    last_move[i] = step
                    n_moved += 1
            if alive[i]:
                assert np.array_equal(w.stored(i), stored[i])
        assert moved == n_moved
        if step == 0:
            assert len(keys) == 0 and moved == 0
        want = _stateless_pairs(stored, alive, last_move, create)
        assert np.array_equal(keys, want), f"step {step}"
        total_pairs += len(keys)
        pos = pos + disp * dyn[:, None]
    assert total_pairs > 0


def test_contact_points_are_project_to_local():
    """contact_point (collision_phases.h:78-82) = particle::project_to_local (core/particle.h:107-108) of the two
    witness points: q* · (p − pos).  Checked against scipy's rotation (1e-12) and for the identity pose (exact)."""
    from scipy.spatial.transform import Rotation

    from scenes import SplitMix64

    rng = SplitMix64(4242)
    n = 64
    pos = rng.uniform(-3.0, 3.0, 2 * n, 3)
    quat = rng.quats(2 * n)
    quat[0] = quat[1] = [0.0, 0.0, 0.0, 1.0]
    pa = np.arange(0, 2 * n, 2, dtype=np.uint32)
    pb = pa + 1
    c10 = rng.uniform(-2.0, 2.0, n, 10)
    out = oracle.contact_points(pos, quat, pa, pb, c10)
    assert np.array_equal(out[0, :3], c10[0, 3:6] - pos[0]) and np.array_equal(out[0, 3:], c10[0, 6:9] - pos[1])
    la = Rotation.from_quat(quat[pa]).inv().apply(c10[:, 3:6] - pos[pa])
    lb = Rotation.from_quat(quat[pb]).inv().apply(c10[:, 6:9] - pos[pb])
    assert np.allclose(out[:, :3], la, atol=1e-12) and np.allclose(out[:, 3:], lb, atol=1e-12)
