"""Pins the oracle's math layer against the reference's own known-answer tests
(math/vector_test.go, quaternion_test.go, matrix_test.go — SURVEY §4 / §8c)."""
import pytest

from math_kat import ALL_KATS
from oracle_lib import Oracle


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("kat", ALL_KATS, ids=lambda f: f.__name__)
def test_oracle_math_kat(kat, prec):
    kat(Oracle(prec))
