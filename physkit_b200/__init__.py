"""physkit_b200 — B200-native collision stage for PhysKit (broadphase LBVH + batched GJK/EPA).

The product is the CUDA library physkit_b200/libpk_collide.so behind the C ABI of
include/pk_collide.h.  This package is the thin Python host mirror used by tests and bench.py; it has
no CPU implementation of anything and raises if the library or a CUDA device is missing.
"""
from .api import (  # noqa: F401
    MODE_QUERY,
    PK_E_EPA_OVERFLOW,
    PK_E_PAIR_OVERFLOW,
    MODE_WORLD,
    RAY_ALL,
    RAY_CLOSEST,
    CollisionWorld,
    Context,
    MultiContext,
    PkError,
    contact_dtype,
    distance_dtype,
    ray_hit_dtype,
    solver_point_dtype,
    gjk_epa,
    library_path,
    load_library,
)
