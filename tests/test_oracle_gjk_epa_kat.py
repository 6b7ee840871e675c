"""Pin the CPU oracle against the reference's own GJK / EPA known-answer tests
(tests/gjk/gjk_test.cpp, tests/epa/epa_test.cpp re-expressed in kat_cases.py)."""
import pytest

import oracle
from kat_cases import EPA_CASES, GJK_CASES, run_epa_case


def _gjk_epa(a, b):
    return oracle.gjk_epa(a[0], (a[1], a[2]), b[0], (b[1], b[2]))


@pytest.mark.parametrize("case", GJK_CASES, ids=[c[0] for c in GJK_CASES])
def test_gjk_kat(case):
    _, a, b, expect, swapped = case
    assert (_gjk_epa(a, b) is not None) == expect
    if swapped:
        assert (_gjk_epa(b, a) is not None) == expect


@pytest.mark.parametrize("case", EPA_CASES, ids=[c[0] for c in EPA_CASES])
def test_epa_kat(case):
    _, a, b, chk = case
    run_epa_case(_gjk_epa, a, b, chk)
