"""GPU: K2 — the hand-written radix sort, the candidate-pair generator against brute force, and
the single large world stepped through the broadphase against the oracle's all-pairs loop (it
must emit the same contacts in the same order)."""
import ctypes as C

import numpy as np
import pytest

from cubez_b200 import _abi, scenes
from golden_cases import STATE_FIELDS
from oracle_lib import OracleWorld

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from cubez_b200.api import Context
    return Context.get(0, "f64")


@pytest.mark.parametrize("n", [1, 2, 31, 257, 2048, 2049, 100_003, 3_000_000])
def test_radix_sort_u32(ctx, n):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    if n > 1000:
        keys[: n // 3] &= 0xFF          # heavy duplicates: stability matters
    vals = np.arange(n, dtype=np.uint32)
    order = np.argsort(keys, kind="stable")
    k, v = keys.copy(), vals.copy()
    ctx.check(ctx.lib.cz_sort_pairs_u32(ctx.h, n, k.ctypes.data_as(C.POINTER(C.c_uint32)), v.ctypes.data_as(C.POINTER(C.c_uint32))))
    assert np.array_equal(k, keys[order]) and np.array_equal(v, vals[order])     # sorted AND stable


@pytest.mark.parametrize("n,bits", [(5, 64), (4097, 40), (300_000, 27)])
def test_radix_sort_u64(ctx, n, bits):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2 ** min(bits, 63), n, dtype=np.uint64)
    vals = rng.integers(0, 2 ** 31, n, dtype=np.uint64).astype(np.uint32)
    order = np.argsort(keys, kind="stable")
    k, v = keys.copy(), vals.copy()
    ctx.check(ctx.lib.cz_sort_pairs_u64(ctx.h, n, k.ctypes.data_as(C.POINTER(C.c_uint64)), v.ctypes.data_as(C.POINTER(C.c_uint32)), bits))
    assert np.array_equal(k, keys[order]) and np.array_equal(v, vals[order])


def brute_pairs(c, r, margin, slack=0.0):
    d2 = ((c[:, None, :] - c[None, :, :]) ** 2).sum(axis=2)
    rr = (r[:, None] + r[None, :]) * margin + slack
    i, j = np.nonzero(np.triu(d2 <= rr * rr, k=1))
    return set(zip(i.tolist(), j.tolist()))


@pytest.fixture(params=["count", "radix"])
def bp_sort(request, monkeypatch):
    """Both sorts by cell key: the one-pass counting sort (default) and the LSD radix sort."""
    monkeypatch.setenv("CUBEZ_BP_SORT", request.param)
    return request.param


@pytest.mark.parametrize("n,spread,seed", [(2, 1.0, 0), (300, 6.0, 1), (1500, 12.0, 2), (1500, 200.0, 3), (4000, 3.0, 4), (4000, 20.0, 5),
                                           (3000, 5000.0, 6)])
def test_candidate_pairs_vs_brute_force(ctx, bp_sort, n, spread, seed):
    """Never drops an overlapping pair (superset of the exact overlaps), never reports a pair
    twice, and reports nothing beyond the documented inflation: 0.5 % on the radii, plus (counting
    path, float sweep) 16 float roundings of a grid-relative coordinate."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-spread, spread, (n, 3))
    if seed == 4:
        c[:, 1] = 0.25           # degenerate axis: a flat layer
    r = rng.uniform(0.2, 1.0, n)
    if seed == 5:
        r[::7] = 0.0             # points
        c[1::2] = c[0::2]        # coincident centres
    cap = min(n * n, 4_000_000)
    pairs = np.zeros((cap, 2), dtype=np.int32)
    cnt = C.c_int64()
    PR = C.POINTER(C.c_double)
    ctx.check(ctx.lib.cz_broadphase_pairs(ctx.h, n, c.ctypes.data_as(PR), r.ctypes.data_as(PR), cap, pairs.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(cnt)))
    got = [tuple(sorted(p)) for p in pairs[: cnt.value].tolist()]
    assert len(got) == len(set(got)), "duplicate candidate pairs"
    got = set(got)
    assert brute_pairs(c, r, 1.0) <= got                       # nothing dropped
    extent = (c.max(axis=0) - c.min(axis=0)).max()
    assert got <= brute_pairs(c, r, 1.0051, 16 * extent * 2.0 ** -24)    # only the documented inflation


@pytest.mark.parametrize("side,frames", [(4, 150), (6, 120)])
def test_pile_world_through_broadphase_matches_all_pairs_oracle(bp_sort, side, frames):
    """cfg3 shape: jittered lattice of cubes and spheres falling into a pile.  The broadphase path
    must produce the contact sequence of the reference's O(n^2) loop, frame by frame."""
    from cubez_b200.api import BatchedWorld
    scene = scenes.pile(side=side)
    gpu = BatchedWorld.from_scene(scene, flags=_abi.WORLD_BROADPHASE)
    cpu = OracleWorld.from_scene(scene)
    for s in range(0, frames, 10):
        gs, cs = gpu.step(scene.dt, 10), cpu.step(scene.dt, 10)
        assert gs["contacts"] == cs["contacts"] and gs["pos_iterations"] == cs["pos_iterations"] and gs["vel_iterations"] == cs["vel_iterations"], s
        assert gpu.contact_pairs(0) == cpu.contact_pairs(0), s
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS:
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    gpu.close()


def test_broadphase_flag_needs_single_all_pairs_world():
    from cubez_b200._abi import CubezError
    from cubez_b200.api import BatchedWorld
    with pytest.raises(CubezError):
        BatchedWorld.from_scene(scenes.batched_cubedrop(n_worlds=2), flags=_abi.WORLD_BROADPHASE)


def test_large_world_resolver_variants_make_identical_choices(monkeypatch):
    """One large world: the three CTA-wide loops — plain (global scans), staged hot values + cached arg-max, and
    adjacency lists + shared caches + prefetched propagation (default) — must make exactly the same choices: same
    iteration counts every frame, bit-identical state.  (Parity with the oracle / the reference at this size is
    tests/test_gpu_cfg3.py and tests/test_gpu_vs_reference_dump.py.)"""
    from cubez_b200.api import BatchedWorld
    scene = scenes.pile(side=8)
    runs = []
    for mode in ("3", "4", "2", "0"):       # 4: mode 3 with the adjacency offsets in global memory as well (what > 23 k contacts use)
        monkeypatch.setenv("CUBEZ_RESOLVE_MODE", mode)
        gpu = BatchedWorld.from_scene(scene, flags=_abi.WORLD_BROADPHASE)
        counts = []
        for _ in range(12):
            st = gpu.step(scene.dt, 5)
            counts.append((st["contacts"], st["pos_iterations"], st["vel_iterations"]))
        runs.append((counts, gpu.checksum_energy()[0], gpu.download()))
        gpu.close()
    assert max(c[2] for c in runs[0][0]) > 1000          # the loops did run
    for other in runs[1:]:
        assert runs[0][0] == other[0]
        assert runs[0][1] == other[1]
        for f in STATE_FIELDS:
            assert np.array_equal(getattr(runs[0][2], f), getattr(other[2], f)), f
