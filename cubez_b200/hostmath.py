"""Host-side (setup-only) math of the cubez `math` package.

The reference keeps mass / inertia setters and scene construction on the host
(SURVEY §8b "What is *not* behind the boundary"); this module mirrors just those few ops in
numpy scalars of the chosen Real type so that setup arithmetic rounds exactly as Go's does
(one IEEE rounding per operation, left to right).  It is not a compute path: per-step work
goes through libcubezcuda only.
"""
from __future__ import annotations

import numpy as np

EPSILON = 1e-7
MIN_NORMAL = 1.1754943508222875e-38


def real_equal(a, b, dtype=np.float64) -> bool:
    """math/math.go:64-78"""
    a, b = dtype(a), dtype(b)
    if a == b:
        return True
    with np.errstate(all="ignore"):
        diff = dtype(abs(dtype(a - b)))
        eps = dtype(EPSILON)
        if dtype(a * b) == 0 or diff < dtype(MIN_NORMAL):
            return bool(diff < dtype(eps * eps))
        return bool(dtype(diff / dtype(np.float64(abs(a)) + np.float64(abs(b)))) < eps)


def m3_determinant(m, dtype=np.float64):
    """math/matrix.go:127-129"""
    m = [dtype(x) for x in m]
    return (m[0] * m[4] * m[8] + m[3] * m[7] * m[2] + m[6] * m[1] * m[5] - m[6] * m[4] * m[2]
            - m[3] * m[1] * m[8] - m[0] * m[7] * m[5])


def m3_invert(m, dtype=np.float64):
    """math/matrix.go:133-153"""
    m = [dtype(x) for x in m]
    with np.errstate(all="ignore"):
        det = m3_determinant(m, dtype)
        if real_equal(det, 0.0, dtype):
            return np.zeros(9, dtype=dtype)
        r = [m[4] * m[8] - m[5] * m[7], m[2] * m[7] - m[1] * m[8], m[1] * m[5] - m[2] * m[4],
             m[5] * m[6] - m[3] * m[8], m[0] * m[8] - m[2] * m[6], m[2] * m[3] - m[0] * m[5],
             m[3] * m[7] - m[4] * m[6], m[1] * m[6] - m[0] * m[7], m[0] * m[4] - m[1] * m[3]]
        s = dtype(1) / det
        return np.array([x * s for x in r], dtype=dtype)


def inertia_tensor_coeffs(ix, iy, iz, ixy=0.0, ixz=0.0, iyz=0.0, dtype=np.float64):
    """math/matrix.go:59-63 (column-major)"""
    ix, iy, iz, ixy, ixz, iyz = (dtype(v) for v in (ix, iy, iz, ixy, ixz, iyz))
    m = np.zeros(9, dtype=dtype)
    m[0], m[3], m[6] = ix, -ixy, -ixz
    m[1], m[4], m[7] = -ixy, iy, -iyz
    m[2], m[5], m[8] = -ixz, -iyz, iz
    return m


def block_inertia_tensor(half, mass, dtype=np.float64):
    """math/matrix.go:68-77"""
    h = [dtype(x) for x in half]
    sq = [h[0] * h[0], h[1] * h[1], h[2] * h[2]]
    mass = dtype(mass)
    c = dtype(0.3)
    return inertia_tensor_coeffs(c * mass * (sq[1] + sq[2]), c * mass * (sq[0] + sq[2]), c * mass * (sq[0] + sq[1]),
                                 0.0, 0.0, 0.0, dtype)


# --- splitmix64 (SURVEY §8d "Random numbers") -------------------------------------------
_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64_draws(seeds: np.ndarray, n_draws: int) -> np.ndarray:
    """For each 64-bit seed return n_draws successive uniforms in [0,1) as float64.
    state += golden; z = state; z = (z ^ z>>30)*M1; z = (z ^ z>>27)*M2; z ^= z>>31;
    u = (z >> 11) * 2^-53."""
    state = np.asarray(seeds, dtype=np.uint64).copy()
    out = np.empty(state.shape + (n_draws,), dtype=np.float64)
    with np.errstate(over="ignore"):
        for k in range(n_draws):
            state = state + _GOLDEN
            z = state.copy()
            z = (z ^ (z >> np.uint64(30))) * _M1
            z = (z ^ (z >> np.uint64(27))) * _M2
            z = z ^ (z >> np.uint64(31))
            out[..., k] = (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)
    return out


def uniform(u, a, b):
    """U(a,b) = a + (b-a)*u evaluated in float64."""
    return np.float64(a) + (np.float64(b) - np.float64(a)) * u
