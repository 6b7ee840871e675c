"""The reference's GJK / EPA known-answer tests re-expressed as data.

Source: /root/reference/tests/gjk/gjk_test.cpp (53 cases) and tests/epa/epa_test.cpp (55 cases);
each entry cites the line range of the case it restates.  The same tables drive the oracle pinning
tests (CPU) and the CUDA parity tests (GPU, through pk_gjk_epa_batch).

A "shape" here is (spec, pos, quat_xyzw):
    aabb(min,max) | obb(center, quat, half) | mbox(half).at(pos, quat) | msphere(r) | mpyr(b,h)
"""
from __future__ import annotations

import math

import numpy as np

from scenes import IDENT, angle_axis, box_vertices, normalized, pyramid_vertices, sphere_vertices

PI = math.pi
Z = (0.0, 0.0, 0.0)


def aabb(mn, mx):
    return (("aabb", tuple(map(float, mn)), tuple(map(float, mx))), Z, IDENT)


def obb(center, quat, half):
    return (("obb", tuple(map(float, half))), tuple(map(float, center)), tuple(quat))


class _Mesh:
    def __init__(self, verts):
        self.verts = verts

    def at(self, pos, quat=IDENT):
        return (("hull", self.verts), tuple(map(float, pos)), tuple(quat))


def mbox(half):
    return _Mesh(box_vertices(half))


_SPH = {}


def msphere(r):
    if r not in _SPH:
        _SPH[r] = _Mesh(sphere_vertices(r))
    return _SPH[r]


def mpyr(b, h):
    return _Mesh(pyramid_vertices(b, h))


UNIT = aabb((-1, -1, -1), (1, 1, 1))
ROT45Z = angle_axis(PI / 4.0, (0, 0, 1))
ROT90Z = angle_axis(PI / 2.0, (0, 0, 1))
ROT90X = angle_axis(PI / 2.0, (1, 0, 0))
ROT180Z = angle_axis(PI, (0, 0, 1))
ROT30_110 = angle_axis(PI / 6.0, normalized((1.0, 1.0, 0.0)))
ROT30_111 = angle_axis(PI / 6.0, normalized((1.0, 1.0, 1.0)))
ROT60_123 = angle_axis(PI / 3.0, normalized((1.0, 2.0, 3.0)))
O1 = obb(Z, IDENT, (1, 1, 1))

# --------------------------------------------------------------------------- GJK boolean KATs
# (name, A, B, expected has_value, also_check_swapped)
GJK_CASES = [
    # AABB vs AABB — gjk_test.cpp:20-89
    ("aabb_aabb_overlap_on_x", UNIT, aabb((0.5, -1, -1), (2.5, 1, 1)), True, False),
    ("aabb_aabb_no_collision_x", UNIT, aabb((2, -1, -1), (4, 1, 1)), False, False),
    ("aabb_aabb_no_collision_y", UNIT, aabb((-1, 2, -1), (1, 4, 1)), False, False),
    ("aabb_aabb_no_collision_z", UNIT, aabb((-1, -1, 2), (1, 1, 4)), False, False),
    ("aabb_aabb_containment", aabb((-3, -3, -3), (3, 3, 3)), UNIT, True, True),
    ("aabb_aabb_large_separation", UNIT, aabb((100, -1, -1), (102, 1, 1)), False, False),
    ("aabb_aabb_same_box", UNIT, UNIT, True, False),
    ("aabb_aabb_partial_overlap_all_axes", aabb((0, 0, 0), (2, 2, 2)), aabb((1, 1, 1), (3, 3, 3)), True, False),
    ("aabb_aabb_separated_diagonal", UNIT, aabb((5, 5, 5), (7, 7, 7)), False, False),
    # OBB vs OBB — gjk_test.cpp:95-207
    ("obb_obb_axis_aligned_overlap", O1, obb((1.5, 0, 0), IDENT, (1, 1, 1)), True, False),
    ("obb_obb_axis_aligned_separated", O1, obb((3, 0, 0), IDENT, (1, 1, 1)), False, False),
    ("obb_obb_same_center_different_orientation", O1, obb(Z, ROT45Z, (1, 1, 1)), True, False),
    ("obb_obb_rotated_45_overlap", O1, obb((1, 0, 0), ROT45Z, (1, 1, 1)), True, False),
    ("obb_obb_rotated_45_separated", O1, obb((3, 0, 0), ROT45Z, (1, 1, 1)), False, False),
    ("obb_obb_cross_config_overlap", obb(Z, IDENT, (2, 0.2, 0.2)), obb(Z, ROT90Z, (2, 0.2, 0.2)), True, False),
    ("obb_obb_cross_config_separated", obb(Z, IDENT, (2, 0.2, 0.2)), obb((0, 3, 0), ROT90Z, (2, 0.2, 0.2)), False, False),
    ("obb_obb_rotated_90_x_overlap", O1, obb((0.5, 0.5, 0.5), ROT90X, (1, 1, 1)), True, False),
    ("obb_obb_non_uniform_extents_overlap", obb(Z, IDENT, (3, 0.3, 3)), obb(Z, IDENT, (0.3, 3, 0.3)), True, False),
    ("obb_obb_non_uniform_extents_separated", obb(Z, IDENT, (3, 0.3, 3)), obb((5, 0, 0), IDENT, (0.3, 3, 0.3)), False, False),
    ("obb_obb_3d_diagonal_overlap", O1, obb((1.5, 1.5, 1.5), IDENT, (1, 1, 1)), True, False),
    ("obb_obb_3d_diagonal_separated", O1, obb((3, 3, 3), IDENT, (1, 1, 1)), False, False),
    # OBB vs AABB — gjk_test.cpp:213-258 (each checks both argument orders)
    ("obb_aabb_axis_aligned_overlap", O1, aabb((0.5, -1, -1), (2.5, 1, 1)), True, True),
    ("obb_aabb_axis_aligned_separated", O1, aabb((2.5, -1, -1), (4.5, 1, 1)), False, True),
    ("obb_aabb_rotated_overlap", obb((1, 0, 0), ROT45Z, (1, 1, 1)), UNIT, True, True),
    ("obb_aabb_rotated_separated", obb((5, 0, 0), ROT45Z, (1, 1, 1)), UNIT, False, True),
    ("obb_aabb_containment", obb(Z, IDENT, (0.5, 0.5, 0.5)), aabb((-2, -2, -2), (2, 2, 2)), True, True),
    # Symmetry — gjk_test.cpp:264-300 (expected value is what both orders must return)
    ("symmetry_aabb_aabb_colliding", UNIT, aabb((0.5, -1, -1), (2.5, 1, 1)), True, True),
    ("symmetry_aabb_aabb_separated", UNIT, aabb((3, -1, -1), (5, 1, 1)), False, True),
    ("symmetry_obb_obb_colliding", O1, obb((1, 1, 0), ROT30_110, (1, 1, 1)), True, True),
    ("symmetry_obb_obb_separated", O1, obb((4, 0, 0), IDENT, (1, 1, 1)), False, True),
    ("collision_info_nullopt_when_separated", UNIT, aabb((3, -1, -1), (5, 1, 1)), False, False),
    # mesh::instance — gjk_test.cpp:306-427
    ("mesh_instance_box_box_overlap", mbox((1, 1, 1)).at(Z), mbox((1, 1, 1)).at((1.5, 0, 0)), True, False),
    ("mesh_instance_box_box_separated", mbox((1, 1, 1)).at(Z), mbox((1, 1, 1)).at((3, 0, 0)), False, False),
    ("mesh_instance_box_box_containment", mbox((3, 3, 3)).at(Z), mbox((0.5, 0.5, 0.5)).at(Z), True, True),
    ("mesh_instance_box_box_same_instance", mbox((1, 1, 1)).at(Z), mbox((1, 1, 1)).at(Z), True, False),
    ("mesh_instance_box_box_rotated_overlap", mbox((1, 1, 1)).at(Z), mbox((1, 1, 1)).at((1, 0, 0), ROT45Z), True, False),
    ("mesh_instance_box_box_rotated_separated", mbox((1, 1, 1)).at(Z), mbox((1, 1, 1)).at((3, 0, 0), ROT45Z), False, False),
    ("mesh_instance_sphere_sphere_overlap", msphere(1.0).at(Z), msphere(1.0).at((1.5, 0, 0)), True, False),
    ("mesh_instance_sphere_sphere_separated", msphere(1.0).at(Z), msphere(1.0).at((3, 0, 0)), False, False),
    ("mesh_instance_box_sphere_overlap", mbox((1, 1, 1)).at(Z), msphere(1.0).at((1.5, 0, 0)), True, True),
    ("mesh_instance_box_sphere_separated", mbox((1, 1, 1)).at(Z), msphere(1.0).at((3, 0, 0)), False, True),
    ("mesh_instance_symmetry_colliding", mbox((1, 1, 1)).at(Z), mbox((1, 1, 1)).at((1.5, 0, 0)), True, True),
    ("mesh_instance_symmetry_separated", mbox((1, 1, 1)).at(Z), mbox((1, 1, 1)).at((4, 0, 0)), False, True),
    ("mesh_instance_collision_info_nullopt_when_separated", mbox((1, 1, 1)).at(Z), mbox((1, 1, 1)).at((4, 0, 0)), False, False),
    # mesh::pyramid — gjk_test.cpp:434-533
    ("pyramid_pyramid_same_pos", mpyr(1, 2).at(Z), mpyr(1, 2).at(Z), True, False),
    ("pyramid_pyramid_overlap_y", mpyr(1, 2).at(Z), mpyr(1, 2).at((0, 1.5, 0)), True, False),
    ("pyramid_pyramid_separated_y", mpyr(1, 2).at(Z), mpyr(1, 2).at((0, 4, 0)), False, False),
    ("pyramid_pyramid_separated_x", mpyr(1, 2).at(Z), mpyr(1, 2).at((4, 0, 0)), False, False),
    ("pyramid_pyramid_flipped_overlap", mpyr(1, 2).at(Z), mpyr(1, 2).at((0, 3, 0), ROT180Z), True, False),
    ("pyramid_pyramid_flipped_separated", mpyr(1, 2).at(Z), mpyr(1, 2).at((0, 6, 0), ROT180Z), False, False),
    ("pyramid_box_overlap", mpyr(1, 2).at(Z), mbox((1, 1, 1)).at(Z), True, True),
    ("pyramid_box_separated", mpyr(1, 2).at(Z), mbox((1, 1, 1)).at((0, -3, 0)), False, True),
    ("pyramid_sphere_overlap", mpyr(1, 2).at(Z), msphere(1.0).at((0, 1, 0)), True, True),
    ("pyramid_sphere_separated", mpyr(1, 2).at(Z), msphere(1.0).at((0, 5, 0)), False, True),
]
# 53 cases registered in main() + collision_info_nullopt_when_separated (defined at :294, unregistered)
assert len(GJK_CASES) == 54

# --------------------------------------------------------------------------- EPA KATs
DEPTH_TOL = 1e-4  # epa_test.cpp:17
MESH_TOL = 0.15  # epa_test.cpp:19
SYM_TOL = 0.05  # epa_test.cpp:21

# Checks understood by run_epa_case():
#   depth=(value, tol)      CHECK_APPROX(result->depth, value, tol)
#   unit=True               check_unit_normal (|n| = 1 ± 1e-6)          epa_test.cpp:24-28
#   axis=(k, lo)            |normal[k]| > lo and the other two < 0.1     epa_test.cpp:164-196
#   axis_min=(k, lo)        |normal[k]| > lo only
#   positive=True           depth > 0
#   mtv=True                moving A by normal*(depth+1e-3) separates    epa_test.cpp:33-47
#   sym_depth=tol           |depth(A,B) - depth(B,A)| <= tol
#   sym_normal=lim          |n(A,B) + n(B,A)| < lim
#   both_depth=(value,tol)  depth of (A,B) and of (B,A) both ≈ value
#   optional=True           result may be nullopt; checks only apply if it has a value
EPA_CASES = [
    # Depth accuracy AABB — epa_test.cpp:57-158
    ("epa_aabb_depth_overlap_x", UNIT, aabb((0.5, -1, -1), (2.5, 1, 1)), dict(depth=(0.5, DEPTH_TOL), unit=True)),
    ("epa_aabb_depth_overlap_y", UNIT, aabb((-1, 0.5, -1), (1, 2.5, 1)), dict(depth=(0.5, DEPTH_TOL), unit=True)),
    ("epa_aabb_depth_overlap_z", UNIT, aabb((-1, -1, 0.5), (1, 1, 2.5)), dict(depth=(0.5, DEPTH_TOL), unit=True)),
    ("epa_aabb_depth_small_overlap", UNIT, aabb((0.9, -1, -1), (2.9, 1, 1)), dict(depth=(0.1, DEPTH_TOL), unit=True)),
    ("epa_aabb_depth_large_overlap", UNIT, aabb((-0.5, -1, -1), (1.5, 1, 1)), dict(depth=(1.5, DEPTH_TOL), unit=True)),
    ("epa_aabb_depth_containment", aabb((-2, -2, -2), (2, 2, 2)), aabb((-0.5, -0.5, -0.5), (0.5, 0.5, 0.5)), dict(depth=(2.5, DEPTH_TOL), unit=True)),
    ("epa_aabb_depth_identical", UNIT, UNIT, dict(depth=(2.0, DEPTH_TOL), unit=True)),
    ("epa_aabb_depth_asymmetric", UNIT, aabb((-0.5, -1, -1), (0.5, 1, 1)), dict(depth=(1.5, DEPTH_TOL), unit=True)),
    ("epa_aabb_depth_corner_overlap", aabb((0, 0, 0), (2, 2, 2)), aabb((1, 1, 1), (3, 3, 3)), dict(depth=(1.0, DEPTH_TOL), unit=True)),
    # Normal direction AABB — epa_test.cpp:164-205
    ("epa_aabb_normal_direction_x", UNIT, aabb((0.5, -1, -1), (2.5, 1, 1)), dict(axis=(0, 0.9))),
    ("epa_aabb_normal_direction_y", UNIT, aabb((-1, 0.5, -1), (1, 2.5, 1)), dict(axis=(1, 0.9))),
    ("epa_aabb_normal_direction_z", UNIT, aabb((-1, -1, 0.5), (1, 1, 2.5)), dict(axis=(2, 0.9))),
    ("epa_aabb_normal_is_unit_length", UNIT, aabb((0.5, 0.3, -0.2), (2.5, 2.3, 1.8)), dict(unit=True)),
    # MTV validity — epa_test.cpp:211-278
    ("epa_mtv_separates_aabb_x", UNIT, aabb((0.5, -1, -1), (2.5, 1, 1)), dict(mtv=True)),
    ("epa_mtv_separates_aabb_diagonal", aabb((0, 0, 0), (2, 2, 2)), aabb((1, 1, 1), (3, 3, 3)), dict(mtv=True)),
    ("epa_mtv_separates_aabb_containment", aabb((-3, -3, -3), (3, 3, 3)), UNIT, dict(mtv=True)),
    ("epa_mtv_separates_obb_axis_aligned", O1, obb((1.5, 0, 0), IDENT, (1, 1, 1)), dict(mtv=True)),
    ("epa_mtv_separates_obb_rotated", O1, obb((1, 0, 0), ROT45Z, (1, 1, 1)), dict(mtv=True)),
    ("epa_mtv_separates_obb_cross", obb(Z, IDENT, (2, 0.2, 0.2)), obb(Z, ROT90Z, (2, 0.2, 0.2)), dict(mtv=True)),
    ("epa_mtv_separates_obb_3d_rotation", O1, obb((1.0, 0.5, 0.3), ROT30_111, (1, 1, 1)), dict(mtv=True)),
    # Symmetry — epa_test.cpp:284-346
    ("epa_symmetry_depth_aabb", UNIT, aabb((0.5, -1, -1), (2.5, 1, 1)), dict(sym_depth=DEPTH_TOL)),
    ("epa_symmetry_normal_aabb", UNIT, aabb((0.5, -1, -1), (2.5, 1, 1)), dict(sym_normal=0.1)),
    ("epa_symmetry_depth_obb", O1, obb((1, 0, 0), ROT45Z, (1, 1, 1)), dict(sym_depth=SYM_TOL)),
    ("epa_symmetry_normal_obb", O1, obb((1, 0, 0), ROT45Z, (1, 1, 1)), dict(sym_normal=0.3)),
    ("epa_symmetry_depth_containment", aabb((-3, -3, -3), (3, 3, 3)), UNIT, dict(sym_depth=DEPTH_TOL)),
    # OBBs — epa_test.cpp:352-425
    ("epa_obb_axis_aligned_depth", O1, obb((1.5, 0, 0), IDENT, (1, 1, 1)), dict(depth=(0.5, DEPTH_TOL), axis_min=(0, 0.9), unit=True)),
    ("epa_obb_identical_at_origin", O1, O1, dict(depth=(2.0, DEPTH_TOL), unit=True)),
    ("epa_obb_cross_config_depth", obb(Z, IDENT, (2, 0.2, 0.2)), obb(Z, ROT90Z, (2, 0.2, 0.2)), dict(depth=(0.4, 1e-2), unit=True)),
    ("epa_obb_rotated_45_depth", O1, obb((1, 0, 0), ROT45Z, (1, 1, 1)), dict(positive=True, unit=True, mtv=True)),
    ("epa_obb_non_uniform_slab_pillar", obb(Z, IDENT, (3, 0.3, 3)), obb(Z, IDENT, (0.3, 3, 0.3)), dict(depth=(3.3, DEPTH_TOL), unit=True)),
    ("epa_obb_rotated_90_x", O1, obb((0.5, 0.5, 0.5), ROT90X, (1, 1, 1)), dict(positive=True, unit=True, mtv=True)),
    # OBB vs AABB — epa_test.cpp:431-453
    ("epa_obb_aabb_axis_aligned_depth", O1, aabb((0.5, -1, -1), (2.5, 1, 1)), dict(both_depth=(0.5, DEPTH_TOL))),
    ("epa_obb_aabb_containment_depth", obb(Z, IDENT, (0.5, 0.5, 0.5)), aabb((-2, -2, -2), (2, 2, 2)), dict(depth=(2.5, DEPTH_TOL), unit=True)),
    # mesh::instance — epa_test.cpp:459-582
    ("epa_mesh_box_box_depth", mbox((1, 1, 1)).at(Z), mbox((1, 1, 1)).at((1.5, 0, 0)), dict(depth=(0.5, DEPTH_TOL), unit=True)),
    ("epa_mesh_box_box_containment_depth", mbox((3, 3, 3)).at(Z), mbox((0.5, 0.5, 0.5)).at(Z), dict(depth=(3.5, DEPTH_TOL), unit=True)),
    ("epa_mesh_box_box_identical_depth", mbox((1, 1, 1)).at(Z), mbox((1, 1, 1)).at(Z), dict(depth=(2.0, DEPTH_TOL), unit=True)),
    ("epa_mesh_box_box_normal_direction", mbox((1, 1, 1)).at(Z), mbox((1, 1, 1)).at((1.5, 0, 0)), dict(axis_min=(0, 0.9))),
    ("epa_mesh_box_box_symmetry", mbox((1, 1, 1)).at(Z), mbox((1, 1, 1)).at((1.5, 0, 0)), dict(sym_depth=DEPTH_TOL)),
    ("epa_mesh_box_box_rotated_mtv", mbox((1, 1, 1)).at(Z), mbox((1, 1, 1)).at((1, 0, 0), ROT45Z), dict(positive=True, unit=True, mtv=True)),
    ("epa_mesh_sphere_sphere_depth", msphere(1.0).at(Z), msphere(1.0).at((1.5, 0, 0)), dict(depth=(0.5, MESH_TOL), unit=True)),
    ("epa_mesh_sphere_sphere_normal", msphere(1.0).at(Z), msphere(1.0).at((1.5, 0, 0)), dict(axis_min=(0, 0.8))),
    ("epa_mesh_sphere_sphere_diagonal", msphere(1.0).at(Z), msphere(1.0).at((1, 1, 0)), dict(depth=(2.0 - math.sqrt(2.0), MESH_TOL), unit=True)),
    ("epa_mesh_box_sphere_depth", mbox((1, 1, 1)).at(Z), msphere(1.0).at((1.5, 0, 0)), dict(depth=(0.5, MESH_TOL), unit=True)),
    # mesh::pyramid — epa_test.cpp:588-638
    ("epa_pyramid_depth_overlap", mpyr(1, 2).at(Z), mpyr(1, 2).at((0, 1.5, 0)), dict(positive=True, unit=True)),
    ("epa_pyramid_mtv_separates", mpyr(1, 2).at(Z), mpyr(1, 2).at((0, 1.5, 0)), dict(mtv=True)),
    ("epa_pyramid_box_mtv", mpyr(1, 2).at(Z), mbox((1, 1, 1)).at(Z), dict(positive=True, unit=True, mtv=True)),
    ("epa_pyramid_flipped_depth", mpyr(1, 2).at(Z), mpyr(1, 2).at((0, 3, 0), ROT180Z), dict(positive=True, unit=True)),
    # Edge cases — epa_test.cpp:685-767
    ("epa_near_touching_aabb", UNIT, aabb((0.99, -1, -1), (2.99, 1, 1)), dict(depth=(0.01, 1e-2), unit=True)),
    ("epa_very_deep_containment", aabb((-100, -100, -100), (100, 100, 100)), aabb((-0.1, -0.1, -0.1), (0.1, 0.1, 0.1)), dict(depth=(100.1, 0.5), unit=True)),
    ("epa_off_center_containment", aabb((-3, -3, -3), (3, 3, 3)), aabb((1, 1, 1), (2, 2, 2)), dict(depth=(2.0, DEPTH_TOL), unit=True)),
    ("epa_flat_slab_overlap", obb(Z, IDENT, (5, 0.1, 5)), obb((0, 0.15, 0), IDENT, (5, 0.1, 5)), dict(depth=(0.05, DEPTH_TOL), axis_min=(1, 0.9), unit=True)),
    ("epa_multiple_rotation_axes", O1, obb((0.5, 0.5, 0.5), ROT60_123, (1, 1, 1)), dict(positive=True, unit=True, mtv=True)),
    ("epa_mixed_mesh_aabb_mtv", mbox((1, 1, 1)).at(Z), aabb((0.5, -1, -1), (2.5, 1, 1)), dict(depth=(0.5, DEPTH_TOL), unit=True)),
    ("epa_mixed_mesh_obb_mtv", mbox((1, 1, 1)).at(Z), obb((1.5, 0, 0), IDENT, (1, 1, 1)), dict(depth=(0.5, DEPTH_TOL), unit=True)),
]
# epa_depth_always_positive (epa_test.cpp:644-679): 7 poses, result optional.
_B = mbox((1, 1, 1))
for _i, (_pb, _rb) in enumerate(
    [
        ((1.5, 0, 0), IDENT),
        ((0, 1.5, 0), IDENT),
        ((0, 0, 1.5), IDENT),
        ((0.5, 0.5, 0.5), IDENT),
        ((0, 0, 0), IDENT),
        ((1.0, 0, 0), ROT45Z),
        ((0.3, 0.3, 0.3), ROT30_111),
    ]
):
    EPA_CASES.append((f"epa_depth_always_positive[{_i}]", _B.at(Z), _B.at(_pb, _rb), dict(optional=True, positive=True, unit=True)))
assert len(EPA_CASES) == 54 + 7  # 54 named cases + the 7-pose sweep of the 55th


def translated(shape, offset):
    """Shape moved by offset (check_mtv_separates, epa_test.cpp:33-47, :528-532)."""
    spec, pos, quat = shape
    off = np.asarray(offset, dtype=np.float64)
    if spec[0] == "aabb":
        mn = np.asarray(spec[1]) + off
        mx = np.asarray(spec[2]) + off
        return (("aabb", tuple(mn), tuple(mx)), pos, quat)
    return (spec, tuple(np.asarray(pos, dtype=np.float64) + off), quat)


def run_epa_case(gjk_epa, a, b, chk):
    """gjk_epa(shapeA, shapeB) -> None | dict(normal, world_a, world_b, depth).  Raises on failure."""
    r = gjk_epa(a, b)
    if chk.get("optional") and r is None:
        return
    assert r is not None, "expected a collision"
    n = np.asarray(r["normal"])
    d = r["depth"]
    if "depth" in chk:
        v, tol = chk["depth"]
        assert abs(d - v) <= tol, (d, v, tol)
    if chk.get("unit"):
        assert abs(math.sqrt(float(n @ n)) - 1.0) <= 1e-6
    if "axis" in chk:
        k, lo = chk["axis"]
        for j in range(3):
            if j == k:
                assert abs(n[j]) > lo
            else:
                assert abs(n[j]) < 0.1
    if "axis_min" in chk:
        k, lo = chk["axis_min"]
        assert abs(n[k]) > lo
    if chk.get("positive"):
        assert d > 0.0
    if chk.get("mtv"):
        moved = translated(a, n * (d + 1e-3))
        assert gjk_epa(moved, b) is None, "MTV does not separate"
    if "sym_depth" in chk or "sym_normal" in chk or "both_depth" in chk:
        r2 = gjk_epa(b, a)
        assert r2 is not None
        if "sym_depth" in chk:
            assert abs(d - r2["depth"]) <= chk["sym_depth"]
        if "sym_normal" in chk:
            s = n + np.asarray(r2["normal"])
            assert math.sqrt(float(s @ s)) < chk["sym_normal"]
        if "both_depth" in chk:
            v, tol = chk["both_depth"]
            assert abs(d - v) <= tol and abs(r2["depth"] - v) <= tol
