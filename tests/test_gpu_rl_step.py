"""GPU: the RL-style step — batched AddVelocity / AddRotation in, observations
out, worlds resident on the device — against the oracle driven the way a host loop over the
reference API would drive it (AddVelocity / AddRotation on every body, then the frame)."""
import numpy as np
import pytest

from cubez_b200 import _abi, scenes
from cubez_b200.api import BatchedWorld, Context
from oracle_lib import OracleWorld

pytestmark = pytest.mark.gpu

OBS = ("position", "orientation", "velocity", "rotation", "motion", "is_awake", "transform", "inverse_inertia_tensor_world",
       "last_frame_acceleration")


def oracle_apply(cpu, av, ar):
    d = cpu.download()
    if av is not None:
        d.velocity[...] = d.velocity + av          # Vector3.Add: one rounding per component
    if ar is not None:
        d.rotation[...] = d.rotation + ar
    cpu.upload_bodies(d)


@pytest.mark.parametrize("n_worlds,flags", [(64, 0), (700, 0), (8, _abi.WORLD_NO_FUSED)])
def test_step_rl_matches_oracle_with_actions(n_worlds, flags):
    sc = scenes.batched_cubedrop(n_worlds=n_worlds)
    ctx = Context.get(0, "f64")
    gpu = BatchedWorld.from_scene(sc, flags=flags, contacts_per_world=64)
    cpu = OracleWorld.from_scene(sc)
    nb = n_worlds * 8
    rng = np.random.default_rng(n_worlds)
    av, ar = ctx.pinned_array((nb, 3)), ctx.pinned_array((nb, 3))
    obs = ctx.pinned_bodies(nb, fields=OBS)
    for frame in range(0, 140, 4):
        kind = (frame // 4) % 4
        a = rng.uniform(-0.3, 0.3, (nb, 3)) * (rng.uniform(0, 1, (nb, 1)) < 0.2) if kind in (0, 1) else None
        r = rng.uniform(-0.5, 0.5, (nb, 3)) * (rng.uniform(0, 1, (nb, 1)) < 0.2) if kind in (1, 2) else None
        if a is not None:
            av[...] = a
        if r is not None:
            ar[...] = r
        gs = gpu.step_rl(av if a is not None else None, ar if r is not None else None, obs, sc.dt, 4)
        oracle_apply(cpu, a, r)
        cs = cpu.step(sc.dt, 4)
        for k in ("contacts", "pos_iterations", "vel_iterations"):
            assert gs[k] == cs[k], (frame, k)
        c = cpu.download()
        for f in OBS:
            assert np.array_equal(getattr(obs, f), getattr(c, f)), (frame, f)
    g = gpu.download()
    for f in OBS:
        assert np.array_equal(getattr(g, f), getattr(obs, f)), f      # the observation is the resident state
    gpu.close()


def test_step_rl_partial_observation_and_no_actions():
    """Only the named arrays are written; without actions the call is cz_world_step + download."""
    sc = scenes.batched_cubedrop(n_worlds=300)
    ctx = Context.get(0, "f64")
    a, b = BatchedWorld.from_scene(sc, contacts_per_world=64), BatchedWorld.from_scene(sc, contacts_per_world=64)
    obs = ctx.pinned_bodies(300 * 8, fields=BatchedWorld.OBS_FIELDS)
    for _ in range(30):
        sa = a.step_rl(None, None, obs, sc.dt, 5)
        sb = b.step(sc.dt, 5)
        assert sa["contacts"] == sb["contacts"] and sa["vel_iterations"] == sb["vel_iterations"]
    d = b.download()
    for f in BatchedWorld.OBS_FIELDS:
        assert np.array_equal(getattr(obs, f), getattr(d, f)), f
    assert obs.transform is None and obs.motion is None
    assert a.checksum_energy()[0] == b.checksum_energy()[0]
    a.close(); b.close()


def test_step_rl_with_episode_resets_f32():
    """Episodes wrap on the device; float32 build."""
    sc = scenes.batched_cubedrop(_abi.F32, n_worlds=96)
    ctx = Context.get(0, "f32")
    gpu = BatchedWorld.from_scene(sc, contacts_per_world=64, ctx=ctx)
    cpu = OracleWorld.from_scene(sc)
    ph = (np.arange(96) * 7 % 50).astype(np.int32)
    gpu.set_episodes(50, ph); cpu.set_episodes(50, ph)
    nb = 96 * 8
    av = ctx.pinned_array((nb, 3))
    obs = ctx.pinned_bodies(nb, fields=BatchedWorld.OBS_FIELDS)
    rng = np.random.default_rng(3)
    for frame in range(0, 120, 3):
        av[...] = rng.uniform(-0.2, 0.2, (nb, 3)).astype(np.float32)
        gpu.step_rl(av, None, obs, sc.dt, 3)
        oracle_apply(cpu, av, None)
        cpu.step(sc.dt, 3)
        c = cpu.download()
        for f in BatchedWorld.OBS_FIELDS:
            assert np.array_equal(getattr(obs, f), getattr(c, f)), (frame, f)
    gpu.close()


def test_step_rl_matches_golden_fixture():
    """cz_world_step_rl against the committed fixture tests/golden/rl12_f64.npz (no oracle at run time)."""
    from golden import make_golden_rl as rl
    from golden_cases import load_golden
    gold = load_golden("rl12_f64")
    sc = rl.make_scene()
    ctx = Context.get(0, "f64")
    gpu = BatchedWorld.from_scene(sc, contacts_per_world=64)
    nb = rl.N_WORLDS * sc.bodies_per_world
    av, ar = ctx.pinned_array((nb, 3)), ctx.pinned_array((nb, 3))
    obs = ctx.pinned_bodies(nb, fields=rl.OBS)
    for call in range(rl.CALLS):
        a, r = rl.actions(call, nb, np.float64)
        if a is not None:
            av[...] = a
        if r is not None:
            ar[...] = r
        st = gpu.step_rl(av if a is not None else None, ar if r is not None else None, obs, sc.dt, rl.FRAMES_PER_CALL)
        assert (st["contacts"], st["pos_iterations"], st["vel_iterations"]) == tuple(int(v) for v in gold["counts"][call]), call
    for f in rl.OBS:
        assert np.array_equal(getattr(obs, f), gold[f]), f
    assert gpu.checksum_energy()[0] == int(gold["checksum"])
    gpu.close()


@pytest.mark.parametrize("n_worlds", [1, 3, 9])
def test_step_rl_tiny_batches_and_zero_frames(n_worlds):
    """Fewer worlds than pipeline chunks; n_steps = 0 applies the actions and returns the observation."""
    sc = scenes.batched_cubedrop(n_worlds=n_worlds)
    ctx = Context.get(0, "f64")
    gpu = BatchedWorld.from_scene(sc, contacts_per_world=64)
    cpu = OracleWorld.from_scene(sc)
    nb = n_worlds * 8
    av = ctx.pinned_array((nb, 3))
    av[...] = np.random.default_rng(n_worlds).uniform(-1, 1, (nb, 3))
    obs = ctx.pinned_bodies(nb, fields=BatchedWorld.OBS_FIELDS)
    gpu.step_rl(av, None, obs, sc.dt, 0)
    oracle_apply(cpu, av, None)
    c = cpu.download()
    for f in BatchedWorld.OBS_FIELDS:
        assert np.array_equal(getattr(obs, f), getattr(c, f)), f
    for _ in range(20):
        gpu.step_rl(av, None, obs, sc.dt, 6)
        oracle_apply(cpu, av, None)
        cpu.step(sc.dt, 6)
    c = cpu.download()
    for f in BatchedWorld.OBS_FIELDS:
        assert np.array_equal(getattr(obs, f), getattr(c, f)), f
    gpu.close()


@pytest.mark.parametrize("n_worlds", [96, 24000])
def test_pipelined_rl_step_equals_the_synchronous_one(n_worlds):
    """cz_world_step_rl_async / cz_world_rl_wait with two steps in flight and alternating buffers: every observation
    (float64 arrays and the float32 ones converted on the device) equals what the synchronous step returns, frame by
    frame; the 24 000-world case runs the split-phase launches."""
    sc = scenes.batched_cubedrop(n_worlds=n_worlds)
    ctx = Context.get(0, "f64")
    a, b = BatchedWorld.from_scene(sc, contacts_per_world=64), BatchedWorld.from_scene(sc, contacts_per_world=64)
    nb = n_worlds * 8
    rng = np.random.default_rng(3)
    frames = 40 if n_worlds < 1000 else 12
    acts = [ctx.pinned_array((nb, 3)) for _ in range(2)]
    obs = [ctx.pinned_bodies(nb, fields=BatchedWorld.OBS_FIELDS) for _ in range(2)]
    obs32 = [{k: ctx.pinned_array((nb, c), dtype=np.float32) for k, c in (("position", 3), ("orientation", 4), ("velocity", 3), ("rotation", 3))} for _ in range(2)]
    ref_act, ref_obs = ctx.pinned_array((nb, 3)), ctx.pinned_bodies(nb, fields=BatchedWorld.OBS_FIELDS)
    actions = [rng.uniform(-0.2, 0.2, (nb, 3)) * (rng.uniform(0, 1, (nb, 1)) < 0.3) for _ in range(frames)]
    expected = []
    for f in range(frames):
        ref_act[...] = actions[f]
        a.step_rl(ref_act, None, ref_obs, sc.dt, 1)
        expected.append({k: getattr(ref_obs, k).copy() for k in BatchedWorld.OBS_FIELDS})
    tickets = []
    def check(f):
        b.rl_wait(tickets[f])
        for k in BatchedWorld.OBS_FIELDS:
            assert np.array_equal(getattr(obs[f & 1], k), expected[f][k]), (f, k)
            assert np.array_equal(obs32[f & 1][k], expected[f][k].astype(np.float32)), (f, k, "f32")
    for f in range(frames):
        if f >= 2:
            check(f - 2)                          # frees slot f & 1
        acts[f & 1][...] = actions[f]
        tickets.append(b.step_rl_async(acts[f & 1], None, obs[f & 1], obs32[f & 1], sc.dt, 1))
    check(frames - 2); check(frames - 1)
    assert a.checksum_energy()[0] == b.checksum_energy()[0]
    from cubez_b200._abi import CubezError
    with pytest.raises(CubezError):
        b.rl_wait(tickets[-1])                    # nothing in flight any more
    a.close(); b.close()


def test_step_rl_with_pageable_host_arrays():
    """Ordinary (pageable) numpy arrays for the actions and the observations: the batched page-locked copy path does not
    apply, the step falls back to one plain copy per field — same values."""
    from cubez_b200._abi import Bodies
    sc = scenes.batched_cubedrop(n_worlds=200)
    gpu = BatchedWorld.from_scene(sc, contacts_per_world=64)
    cpu = OracleWorld.from_scene(sc)
    nb = 200 * 8
    rng = np.random.default_rng(3)
    obs = Bodies(nb, gpu.prec, fields=BatchedWorld.OBS_FIELDS)
    for it, n in enumerate((1, 30, 2, 50)):
        a = np.ascontiguousarray(rng.uniform(-0.3, 0.3, (nb, 3)) * (rng.uniform(0, 1, (nb, 1)) < 0.2))
        gs = gpu.step_rl(a, None, obs, sc.dt, n)
        oracle_apply(cpu, a, None)
        cs = cpu.step(sc.dt, n, n_threads=8)
        for k in ("contacts", "pos_iterations", "vel_iterations"):
            assert gs[k] == cs[k], (it, k)
        c = cpu.download()
        for f in BatchedWorld.OBS_FIELDS:
            assert np.array_equal(getattr(obs, f), getattr(c, f)), (it, f)
    gpu.close()
