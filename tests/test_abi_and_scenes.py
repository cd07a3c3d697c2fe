"""CPU: the C-ABI library loads and exports every symbol include/cubezcuda.h declares (no
compute calls), fails loudly without a device, and the scene builders produce the
documented initial conditions."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from cubez_b200 import _abi, scenes
from cubez_b200.hostmath import block_inertia_tensor, m3_invert, real_equal, splitmix64_draws

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "cubezcuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cz_[a-z0-9_]+)\s*\(", text)))


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_library_exports_every_declared_symbol(prec):
    lib = _abi.load(prec)
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cubezcuda.h but not exported"
    assert sorted("cz_" + s for s in _abi.EXPORTED) == names
    assert lib.cz_real_size() == (8 if prec == "f64" else 4)


def test_no_cpu_fallback_without_device():
    """Without a CUDA device cz_init must fail with CZ_ERR_CUDA — never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _abi.load("f64")
    h = C.c_void_p()
    rc = lib.cz_init(0, C.byref(h))
    assert rc == _abi.CZ_ERR_CUDA
    assert b"no CPU fallback" in lib.cz_last_error(None)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cubez_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in src and "cubez_oracle" not in src and "liboracle" not in src and "hostemu" not in src.replace("tests/hostemu", ""), f


def test_cubedrop_initial_conditions():
    sc = scenes.cubedrop()
    assert np.array_equal(sc.bodies.position[:, 0], [-2.5, -0.5, 1.5, 3.5, -1.75, 0.25, 2.25, 4.25])
    assert np.all(sc.bodies.position[:, 1] == 10.0) and np.all(sc.bodies.position[:, 2] == 0.0)
    assert np.all(sc.bodies.inverse_mass == 0.125)
    it = block_inertia_tensor((0.5, 0.5, 0.5), 8.0)
    assert np.array_equal(sc.bodies.inverse_inertia_tensor[0], m3_invert(it))
    assert real_equal(it[0], 0.3 * 8 * 0.5)


def test_ballistic_schedule_matches_reference_loop_order():
    sc = scenes.ballistic(n_bullets=3)
    # examples/ballistic.go:47-97 — (cube,plane),(cube,backboard), per bullet: (b,plane),(cube,b),(backboard,b),(b2,b)...
    want = [(0, -1), (0, 1),
            (2, -1), (0, 2), (1, 2), (3, 2), (4, 2),
            (3, -1), (0, 3), (1, 3), (2, 3), (4, 3),
            (4, -1), (0, 4), (1, 4), (2, 4), (3, 4)]
    assert list(zip(sc.check_one.tolist(), sc.check_two.tolist())) == want
    assert sc.integrate.tolist() == [1, 0, 1, 1, 1] and sc.active_from.tolist() == [0, 0, 60, 68, 76]
    full = scenes.ballistic()
    assert full.check_one.shape[0] == 2 + 64 * (3 + 63)     # P <= 4 226


def test_splitmix64_known_values_and_shard_invariance():
    # first outputs of splitmix64(seed=0): 0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4
    u = splitmix64_draws(np.array([0], dtype=np.uint64), 2)[0]
    assert u[0] == (0xE220A8397B1DCDAF >> 11) * 2.0 ** -53 and u[1] == (0x6E789E6AA1B965F4 >> 11) * 2.0 ** -53
    whole = scenes.batched_cubedrop(n_worlds=16)
    part = scenes.batched_cubedrop(n_worlds=8, first_world=8)
    assert np.array_equal(whole.bodies.position[64:], part.bodies.position)
    assert np.array_equal(whole.bodies.orientation[64:], part.bodies.orientation)
    q = whole.bodies.orientation
    assert np.allclose((q * q).sum(axis=1), 1.0, atol=1e-15)


def test_pile_has_no_initial_overlap():
    sc = scenes.pile(side=6)
    p = sc.bodies.position
    d = np.linalg.norm(p[:, None, :] - p[None, :, :], axis=2) + np.eye(p.shape[0]) * 10
    assert d.min() > 1.0 and sc.bodies_per_world == 216
    assert (sc.colliders.shape == _abi.SHAPE_CUBE).sum() == 108
