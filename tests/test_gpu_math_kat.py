"""The same 22 reference known-answer tests, run on the device math of libcubezcuda, plus a
bit-exact differential of every math op against the oracle on random operands."""
import numpy as np
import pytest

from cubez_b200 import _abi
from math_kat import ALL_KATS
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["f64", "f32"])
def backends(request):
    from cubez_b200.api import Context
    return Context.get(0, request.param), Oracle(request.param)


@pytest.mark.parametrize("kat", ALL_KATS, ids=lambda f: f.__name__)
def test_device_math_kat(kat, backends):
    kat(backends[0])


OP_IN = {"VEC_ADD": 6, "VEC_ADD_SCALED": 7, "VEC_COMPONENT_PRODUCT": 6, "VEC_CROSS": 6, "VEC_DOT": 6, "VEC_MAGNITUDE": 3,
         "VEC_SQUARE_MAGNITUDE": 3, "VEC_MUL_WITH": 4, "VEC_NORMALIZE": 3, "VEC_SUB": 6, "QUAT_MUL": 8, "QUAT_LEN": 4,
         "QUAT_NORMALIZE": 4, "QUAT_ROTATE": 7, "QUAT_ADD_SCALED_VECTOR": 8, "M3_MUL_M3": 18, "M3_INVERT": 9, "M3_MUL_V": 12,
         "M3_TRANSFORM_TRANSPOSE": 12, "M3_DETERMINANT": 9, "M34_MUL_M34": 24, "M34_MUL_V": 15, "M34_TRANSFORM_INVERSE": 15,
         "M34_SET_AS_TRANSFORM": 7, "REAL_EQUAL": 2, "TRANSFORM_INERTIA": 21}


@pytest.mark.parametrize("op", sorted(OP_IN))
def test_device_math_bit_exact_vs_oracle(op, backends):
    gpu, cpu = backends
    rng = np.random.default_rng(hash(op) % (2 ** 32))
    for trial in range(8):
        x = rng.uniform(-3, 3, OP_IN[op])
        if trial == 0:
            x[:] = 0          # degenerate operands: zero vector / singular matrix / zero quaternion
        if trial == 1 and op == "REAL_EQUAL":
            x[1] = x[0] * (1 + 5e-8)
        a, b = gpu.math_op(op, x), cpu.math_op(op, x)
        assert np.array_equal(a, b, equal_nan=True), (op, x, a, b)
