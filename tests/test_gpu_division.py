"""The library replaces the three IEEE divisions of Eigen's normalized() (reference lin_alg.h:232-240) by
Markstein's division through one correctly rounded reciprocal (pk_common.cuh: pk_div_by_rcp).  Results must
be bit-identical to `/`; the device self-test draws random and adversarial operand pairs (numerators next to
exact multiples of the divisor, all-ones / all-zero divisor mantissas) and counts differing quotients."""
import pytest

import physkit_b200 as pk

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 0x5EED, 20261017])
def test_shared_reciprocal_division_is_bit_identical(seed):
    ctx = pk.Context(max_bodies=16, max_pairs=16, max_shapes=4)
    try:
        assert ctx.selftest_division(seed, 1 << 31) == 0
    finally:
        ctx.close()
