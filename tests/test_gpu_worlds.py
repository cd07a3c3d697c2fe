"""GPU: the batched-world handle vs golden fixtures and vs the oracle — contact sets
(pair sequence, counts) bit-exact per step, state bit-exact, on both the multi-kernel path
and the fused small-world kernel, f64 and f32."""
import os

import numpy as np
import pytest

import hostemu_lib as he
from cubez_b200 import _abi, scenes
from golden_cases import CASES, STATE_FIELDS, check_against_golden, load_golden
from oracle_lib import OracleWorld

pytestmark = pytest.mark.gpu
PATHS = {"multi": (_abi.WORLD_NO_FUSED, None), "fused32": (_abi.WORLD_FUSED, "32"), "fused16": (_abi.WORLD_FUSED, "16"),
         "fused8": (_abi.WORLD_FUSED, "8")}


def make_world(scene, path, **kw):
    from cubez_b200.api import BatchedWorld
    flags, g = PATHS[path]
    if g:
        os.environ["CUBEZ_FUSED_G"] = g
    try:
        return BatchedWorld.from_scene(scene, flags=flags, **kw)
    finally:
        os.environ.pop("CUBEZ_FUSED_G", None)


@pytest.mark.parametrize("path", sorted(PATHS))
@pytest.mark.parametrize("name", sorted(CASES))
def test_world_matches_golden(name, path):
    make, n = CASES[name]
    scene = make()
    w = make_world(scene, path)
    check_against_golden(w, scene, load_golden(name), n, he.pair_hash)
    w.close()


@pytest.mark.parametrize("path", ["multi", "fused8"])
def test_cubedrop_600_steps_vs_oracle(path):
    """cfg1 full length: contact sets compared every step for all 600 steps."""
    scene = scenes.cubedrop()
    gpu, cpu = make_world(scene, path), OracleWorld.from_scene(scene)
    for s in range(600):
        gpu.step(scene.dt, 1); cpu.step(scene.dt, 1)
        assert gpu.contact_pairs(0) == cpu.contact_pairs(0), s
        if s % 50 == 0:
            gc, cc = gpu.contacts(0), cpu.contacts(0)
            for f in ("point", "normal", "penetration"):
                assert np.array_equal(gc.valid(f), cc.valid(f)), (s, f)
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS + ("transform", "inverse_inertia_tensor_world", "last_frame_acceleration"):
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    assert not g.is_awake.any()            # everything has gone to sleep
    gpu.close()


def test_ballistic_full_600_steps_vs_oracle():
    """cfg2: 64 bullets spawned over time, explicit 4 226-check schedule, static backboard.
    (4 226 checks exceed the fused kernel's per-world schedule limit, so this is the multi-kernel
    path: multi-tile count / scan / emit narrowphase + warp resolver; the fused kernel is covered
    on the 16-bullet golden case.)"""
    from cubez_b200._abi import CubezError
    scene = scenes.ballistic()
    forced = make_world(scene, "fused32")
    with pytest.raises(CubezError) as e:          # asking for the fused kernel must fail loudly, not fall back
        forced.step(scene.dt, 1)
    assert e.value.code == _abi.CZ_ERR_INVALID
    forced.close()
    gpu, cpu = make_world(scene, "multi"), OracleWorld.from_scene(scene)
    for s in range(0, 600, 4):
        gs, cs = gpu.step(scene.dt, 4), cpu.step(scene.dt, 4)
        assert gs["contacts"] == cs["contacts"] and gs["vel_iterations"] == cs["vel_iterations"] and gs["pos_iterations"] == cs["pos_iterations"], s
        assert gpu.contact_pairs(0) == cpu.contact_pairs(0), s
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS:
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    assert gpu.checksum_energy()[0] == cpu.checksum_energy()[0]
    gpu.close()


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_batched_worlds_multi_step_calls_and_checksum(prec):
    """cfg4 shape (512 worlds): n_steps per call, parity on every world, checksum equality, and
    shard invariance (two half-size handles == one full handle)."""
    p = _abi.precision(prec)
    scene = scenes.batched_cubedrop(p, n_worlds=512)
    gpu, cpu = make_world(scene, "fused8"), OracleWorld.from_scene(scene)
    for s in range(0, 240, 40):
        gs, cs = gpu.step(scene.dt, 40), cpu.step(scene.dt, 40, n_threads=8)
        for k in ("contacts", "pos_iterations", "vel_iterations", "max_contacts"):
            assert gs[k] == cs[k], (s, k, gs[k], cs[k])
        gc, cc = gpu.last_counts(), cpu.last_counts()
        for a, b in zip(gc, cc):
            assert np.array_equal(a, b)
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS:
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    total = gpu.checksum_energy()
    assert total[0] == cpu.checksum_energy()[0]
    assert abs(total[1] - cpu.checksum_energy()[1]) <= 1e-9 * abs(total[1])
    gpu.close()
    # shards
    acc = 0
    for first in (0, 256):
        sh = scenes.batched_cubedrop(p, n_worlds=256, first_world=first)
        w = make_world(sh, "fused8")
        w.step(sh.dt, 240, stats=False)
        acc = (acc + w.checksum_energy()[0]) % (1 << 64)
        w.close()
    assert acc == total[0]


def test_pile_small_all_pairs_multi_tile():
    """cfg3 shape at 6x6x6 = 216 bodies: 46 872 checks per step -> multi-tile count/scan/emit
    narrowphase and the CTA-wide resolver with global scratch."""
    scene = scenes.pile(side=6)
    gpu, cpu = make_world(scene, "multi"), OracleWorld.from_scene(scene)
    for s in range(0, 120, 10):
        gs, cs = gpu.step(scene.dt, 10), cpu.step(scene.dt, 10)
        assert gs["contacts"] == cs["contacts"] and gs["pos_iterations"] == cs["pos_iterations"] and gs["vel_iterations"] == cs["vel_iterations"], s
        assert gpu.contact_pairs(0) == cpu.contact_pairs(0), s
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS:
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    gpu.close()


def test_contact_capacity_overflow_is_an_error():
    from cubez_b200._abi import CubezError
    scene = scenes.cubedrop()
    for path in ("multi", "fused8"):
        w = make_world(scene, path, contacts_per_world=16)
        with pytest.raises(CubezError) as e:
            w.step(scene.dt, 200)
        assert e.value.code == _abi.CZ_ERR_CAPACITY
        w.close()


def test_upload_download_roundtrip_and_step_host():
    """cz_world_step_host (host buffers in/out) == device-resident stepping."""
    scene = scenes.batched_cubedrop(n_worlds=64)
    a, b = make_world(scene, "fused8"), make_world(scene, "fused8")
    host = a.download()
    for s in range(100):
        a.step_host(host, scene.dt, 1)
        b.step(scene.dt, 1, stats=False)
    ref = b.download()
    for f in STATE_FIELDS + ("transform", "inverse_inertia_tensor_world", "last_frame_acceleration"):
        assert np.array_equal(getattr(host, f), getattr(ref, f)), f
    a.close(); b.close()


def test_set_pow_override_matches_default():
    scene = scenes.cubedrop()
    a, b = make_world(scene, "multi"), make_world(scene, "multi")
    dt = scene.dt
    lp = np.power(scene.bodies.linear_damping.astype(np.float64), dt)
    ap = np.power(scene.bodies.angular_damping.astype(np.float64), dt)
    a.set_pow(dt, lp, ap, 0.5 ** dt)
    a.step(dt, 100); b.step(dt, 100)
    assert a.checksum_energy()[0] == b.checksum_energy()[0]
    a.close(); b.close()


@pytest.mark.parametrize("path", ["multi", "fused8", "fused32"])
def test_episode_reset_staggered_phases(path):
    """RL-style episodes: worlds at staggered phases, restored to the snapshot on wrap."""
    scene = scenes.batched_cubedrop(n_worlds=24)
    gpu, cpu = make_world(scene, path), OracleWorld.from_scene(scene)
    L = 130
    phase0 = (np.arange(24) * 7) % L
    gpu.set_episodes(L, phase0); cpu.set_episodes(L, phase0)
    for s in range(0, 300, 25):
        gs, cs = gpu.step(scene.dt, 25), cpu.step(scene.dt, 25)
        for k in ("contacts", "pos_iterations", "vel_iterations"):
            assert gs[k] == cs[k], (s, k)
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS:
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    gpu.close()


@pytest.mark.parametrize("path", ["multi", "fused8", "fused32"])
def test_collider_offsets_and_two_planes(path):
    """Non-identity collider Offset matrices (colliders.go:43,61) and a second, tilted plane:
    exercises the full transform = body.transform x Offset product and multi-plane schedules."""
    scene = scenes.batched_cubedrop(n_worlds=6)
    rng = np.random.default_rng(11)
    n = scene.bodies.n
    off = scene.colliders.offset
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    off[:, 0] = 1 - 2 * y * y - 2 * z * z; off[:, 1] = 2 * x * y + 2 * w * z; off[:, 2] = 2 * x * z - 2 * w * y
    off[:, 3] = 2 * x * y - 2 * w * z; off[:, 4] = 1 - 2 * x * x - 2 * z * z; off[:, 5] = 2 * y * z + 2 * w * x
    off[:, 6] = 2 * x * z + 2 * w * y; off[:, 7] = 2 * y * z - 2 * w * x; off[:, 8] = 1 - 2 * x * x - 2 * y * y
    off[:, 9:12] = rng.uniform(-0.2, 0.2, (n, 3))
    off[::3] = (1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0)          # every third collider keeps the identity (fast path)
    scene.planes = _abi.Planes([[0, 1, 0], [0.6, 0.8, 0.0]], [0.0, -3.0], scene.prec)
    gpu, cpu = make_world(scene, path), OracleWorld.from_scene(scene)
    for s in range(0, 200, 20):
        gs, cs = gpu.step(scene.dt, 20), cpu.step(scene.dt, 20)
        for k in ("contacts", "pos_iterations", "vel_iterations"):
            assert gs[k] == cs[k], (s, k, gs[k], cs[k])
        assert gpu.contact_pairs(3) == cpu.contact_pairs(3)
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS + ("transform",):
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    assert np.array_equal(gpu.download_colliders().transform, cpu.download_colliders().transform)
    gpu.close()


def test_split_phase_launches_equal_persistent_kernel():
    """Large batches run one launch per phase of the frame (contact state in per-world global
    arrays between launches); it must give exactly the persistent kernel's results."""
    scene = scenes.batched_cubedrop(n_worlds=96)
    phase0 = (np.arange(96) * 5) % 140
    sums = []
    for split in ("0", "1"):
        os.environ["CUBEZ_FUSED_SPLIT"] = split
        try:
            w = make_world(scene, "fused8")
        finally:
            os.environ.pop("CUBEZ_FUSED_SPLIT")
        w.set_episodes(140, phase0)
        st = w.step(scene.dt, 300)
        sums.append((w.checksum_energy()[0], st["contacts"], st["pos_iterations"], st["vel_iterations"], tuple(w.contact_pairs(7))))
        w.close()
    assert sums[0] == sums[1]
    cpu = OracleWorld.from_scene(scene)
    cpu.set_episodes(140, phase0)
    cs = cpu.step(scene.dt, 300, n_threads=8)
    assert sums[1][0] == cpu.checksum_energy()[0] and sums[1][1:4] == (cs["contacts"], cs["pos_iterations"], cs["vel_iterations"])


@pytest.mark.parametrize("lanes,prec", [("2", _abi.F64), ("3", _abi.F64), ("4", _abi.F64), ("3", _abi.F32)])
def test_split_step_in_lanes_matches_the_oracle(lanes, prec, monkeypatch):
    """The resident split-mode step runs the batch as independent slices on their own streams (world counters, group
    scratch and ordering per slice).  Uneven slices, episode resets, several calls: state, counters and the last
    frame's contacts must be the oracle's, exactly as with one lane."""
    scene = scenes.batched_cubedrop(prec, n_worlds=250)
    phase0 = (np.arange(250) * 7) % 150
    monkeypatch.setenv("CUBEZ_FUSED_SPLIT", "1")
    monkeypatch.setenv("CUBEZ_STEP_LANES", lanes)
    monkeypatch.setenv("CUBEZ_STEP_LANE_MIN", "16")
    gpu = make_world(scene, "fused8")
    cpu = OracleWorld.from_scene(scene)
    gpu.set_episodes(150, phase0)
    cpu.set_episodes(150, phase0)
    for n in (1, 60, 3, 140):
        gs, cs = gpu.step(scene.dt, n), cpu.step(scene.dt, n, n_threads=8)
        for k in ("contacts", "pos_iterations", "vel_iterations"):
            assert gs[k] == cs[k], (n, k, gs[k], cs[k])
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS:
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    for wi in (0, 83, 84, 249):   # around the slice edges
        assert gpu.contact_pairs(wi) == cpu.contact_pairs(wi)
    gpu.close()


def test_cost_ordered_scheduling_is_result_neutral(monkeypatch):
    """czf::k_order_worlds only permutes the order in which worlds are fetched: state, counters and
    checksum are identical with and without it, resident and chunked."""
    import numpy as np
    from cubez_b200 import scenes
    from cubez_b200.api import BatchedWorld
    sc = scenes.batched_cubedrop(n_worlds=1500)
    res = []
    for order in ("1", "0"):
        monkeypatch.setenv("CUBEZ_FUSED_ORDER", order)
        monkeypatch.setenv("CUBEZ_FUSED_SPLIT", "1")
        w = BatchedWorld.from_scene(sc, contacts_per_world=64)
        tot = {"contacts": 0, "pos_iterations": 0, "vel_iterations": 0}
        for _ in range(40):
            st = w.step(sc.dt, 5)
            for k in tot:
                tot[k] += st[k]
        for _ in range(10):
            st = w.step_rl(None, None, None, sc.dt, 2)
            for k in tot:
                tot[k] += st[k]
        res.append((tot, w.checksum_energy()[0], w.last_counts()))
        w.close()
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1]
    for a, b in zip(res[0][2], res[1][2]):
        assert np.array_equal(a, b)


def test_contact_capacity_exact_fit_and_one_short():
    """Edge of the capacity contract: a capacity equal to the largest contact count of the run works (and gives the
    golden trajectory); one slot less is CZ_ERR_CAPACITY, never a silent truncation (it would change 8*len)."""
    from cubez_b200._abi import CubezError
    gold = load_golden("cubedrop_f64")
    peak = int(gold["counts"].max())
    scene = scenes.cubedrop()
    for path in ("multi", "fused8"):
        w = make_world(scene, path, contacts_per_world=peak)
        check_against_golden(w, scene, gold, CASES["cubedrop_f64"][1], he.pair_hash)
        w.close()
        w = make_world(scene, path, contacts_per_world=peak - 1)
        with pytest.raises(CubezError) as e:
            w.step(scene.dt, CASES["cubedrop_f64"][1])
        assert e.value.code == _abi.CZ_ERR_CAPACITY
        w.close()


def test_full_size_batch_equals_the_sum_of_its_parts():
    """BASELINE cfg4 at full size (65 536 worlds, split-phase launches) against the same worlds stepped as four
    batches of 16 384 (persistent kernel): worlds are independent, so checksums (sum mod 2^64) and counters add up."""
    from cubez_b200.api import BatchedWorld
    W, parts, frames = 65536, 4, 130
    def run(first, n):
        sc = scenes.batched_cubedrop(n_worlds=n, first_world=first)
        w = BatchedWorld.from_scene(sc, contacts_per_world=64)
        w.set_episodes(600, ((first + np.arange(n)) % 600).astype(np.int32))
        st = w.step(sc.dt, frames)
        out = (w.checksum_energy()[0], st["contacts"], st["pos_iterations"], st["vel_iterations"])
        w.close()
        return out
    whole = run(0, W)
    acc = [0, 0, 0, 0]
    for k in range(parts):
        r = run(k * (W // parts), W // parts)
        acc = [a + b for a, b in zip(acc, r)]
    acc[0] &= (1 << 64) - 1
    assert whole[1] > 0 and whole[3] > 0
    assert tuple(acc) == whole


@pytest.mark.parametrize("split", ["0", "1"])
def test_f32_fused_records_at_exact_capacity(split, monkeypatch):
    """float32 build, persistent and split-phase launches, contact capacity EQUAL to the peak count: the per-body
    contact masks are 64-bit whatever Real is, so the shared-memory record must be sized in bytes (an f32 record sized
    in reals let a world within 4 contacts of capacity overwrite its neighbour in the CTA)."""
    scene = scenes.batched_cubedrop(_abi.F32, n_worlds=80)
    cpu = OracleWorld.from_scene(scene)
    ref = [cpu.step(scene.dt, 20, n_threads=8) for _ in range(10)]
    peak = max(r["max_contacts"] for r in ref)
    monkeypatch.setenv("CUBEZ_FUSED_SPLIT", split)
    gpu = make_world(scene, "fused8", contacts_per_world=peak)
    for r in ref:
        gs = gpu.step(scene.dt, 20)
        for k in ("contacts", "pos_iterations", "vel_iterations", "max_contacts"):
            assert gs[k] == r[k], (k, gs[k], r[k])
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS:
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    gpu.close()


def test_f32_wide_group_record_ballistic():
    """float32, 18 bodies per world (32 lanes per world, persistent kernel): the always-written per-check tails of
    the record follow the 64-bit masks."""
    scene = scenes.ballistic(_abi.F32, n_bullets=16)
    gpu, cpu = make_world(scene, "fused32"), OracleWorld.from_scene(scene)
    for s in range(0, 300, 10):
        gs, cs = gpu.step(scene.dt, 10), cpu.step(scene.dt, 10)
        for k in ("contacts", "pos_iterations", "vel_iterations"):
            assert gs[k] == cs[k], (s, k, gs[k], cs[k])
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS:
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    gpu.close()


def test_replan_after_host_step_rebuilds_the_pipeline():
    """cz_world_upload_planes after a host-buffer step re-plans the world; the chunk pipeline (sized from the old
    plan) must be rebuilt, not reused."""
    scene = scenes.batched_cubedrop(n_worlds=48)
    a, b = make_world(scene, "fused8"), make_world(scene, "fused8")
    host = a.download()
    a.step_host(host, scene.dt, 1); b.step(scene.dt, 1, stats=False)
    two = _abi.Planes([[0, 1, 0], [0.6, 0.8, 0.0]], [0.0, -3.0], scene.prec)
    a.upload_planes(two); b.upload_planes(two)          # more checks per world -> larger records, new scratch sizes
    for _ in range(60):
        a.step_host(host, scene.dt, 1); b.step(scene.dt, 1, stats=False)
    ref = b.download()
    for f in STATE_FIELDS:
        assert np.array_equal(getattr(host, f), getattr(ref, f)), f
    a.close(); b.close()


def test_device_status_is_sticky_across_async_steps():
    """An asynchronous step (stats=NULL) cannot return a device-side error; it must surface at the next observing
    call instead of being wiped by it."""
    from cubez_b200._abi import CubezError
    scene = scenes.cubedrop()
    w = make_world(scene, "fused8", contacts_per_world=16)
    w.step(scene.dt, 200, stats=False)                 # overflows around frame 90; nothing can be reported here
    with pytest.raises(CubezError) as e:
        w.synchronize()
    assert e.value.code == _abi.CZ_ERR_CAPACITY
    w.synchronize()                                    # observed once: cleared
    w.close()


def test_nonfinite_body_count():
    """cz_world_count_nonfinite: the NaN / overflow watch of SURVEY section 5.  Two static cubes in contact make 0/0 = NaN
    moves in applyPositionChange (contact.go:329-330) — silently, in the reference and here; the counter is how a host notices."""
    scene = scenes.cubedrop()
    w = make_world(scene, "multi")
    w.step(scene.dt, 30)
    assert w.count_nonfinite() == 0
    b = w.download()
    b.inverse_mass[:2] = 0.0                                  # two infinite-mass cubes ...
    b.inverse_inertia_tensor[:2] = 0.0
    b.position[0] = (0.0, 5.0, 0.0); b.position[1] = (0.3, 5.2, 0.0)    # ... overlapping
    b.velocity[7] = (np.inf, 0.0, 0.0)                        # and an overflowed velocity
    w.upload_bodies(b, derive=True)
    w.upload_colliders(scene.colliders, derive=True)
    w.step(scene.dt, 2)
    g = w.download()
    expected = int((~np.isfinite(np.concatenate([g.position, g.orientation, g.velocity, g.rotation], axis=1))).any(axis=1).sum())
    assert expected >= 3 and w.count_nonfinite() == expected
    w.close()
