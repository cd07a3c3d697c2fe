"""GPU: random worlds (scenes.random_worlds — cubes, spheres, collider-less bodies, collider Offsets, several planes,
sleeping / non-sleeping bodies, late activation, per-body damping and gravity) stepped on every execution path of the
library against the CPU oracle: counters per block of frames, contact sets of a few worlds, every body's state bits."""
import os

import numpy as np
import pytest

from cubez_b200 import _abi, scenes
from golden_cases import STATE_FIELDS
from oracle_lib import OracleWorld

pytestmark = pytest.mark.gpu

# (bodies per world, worlds, contact capacity, environment of the plan, world flags); capacities on both sides of 64 and
# 256 pick the three propagation variants of the fused loops (per-body bitmasks, match list, plain scan)
PATHS = {
    "fused8": (8, 96, 64, {"CUBEZ_FUSED_G": "8", "CUBEZ_FUSED_SPLIT": "0"}, _abi.WORLD_FUSED),
    "split-lanes": (8, 200, 64, {"CUBEZ_FUSED_G": "8", "CUBEZ_FUSED_SPLIT": "1", "CUBEZ_STEP_LANES": "3", "CUBEZ_STEP_LANE_MIN": "16"}, _abi.WORLD_FUSED),
    "split-matchlist": (8, 150, 192, {"CUBEZ_FUSED_G": "8", "CUBEZ_FUSED_SPLIT": "1", "CUBEZ_STEP_LANES": "2", "CUBEZ_STEP_LANE_MIN": "16"}, _abi.WORLD_FUSED),
    "fused16": (13, 40, 128, {"CUBEZ_FUSED_G": "16"}, _abi.WORLD_FUSED),
    "fused32": (24, 20, 320, {"CUBEZ_FUSED_G": "32"}, _abi.WORLD_FUSED),
    "multi": (10, 24, 240, {}, _abi.WORLD_NO_FUSED),
}


@pytest.mark.parametrize("prec", [_abi.F64, _abi.F32], ids=["f64", "f32"])
@pytest.mark.parametrize("seed", [3, 17, 40, 101])
@pytest.mark.parametrize("path", sorted(PATHS))
def test_random_worlds_match_the_oracle(path, seed, prec, monkeypatch):
    from cubez_b200.api import BatchedWorld
    B, W, cap, env, flags = PATHS[path]
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    scene = scenes.random_worlds(prec, n_worlds=W, bodies_per_world=B, seed=seed, n_planes=1 + seed % 3)
    gpu = BatchedWorld.from_scene(scene, flags=flags, contacts_per_world=cap)
    cpu = OracleWorld.from_scene(scene)
    for block in range(6):
        gs, cs = gpu.step(scene.dt, 20), cpu.step(scene.dt, 20, n_threads=8)
        assert cs["status"] == 0 and cs["max_contacts"] <= cap
        for k in ("contacts", "pos_iterations", "vel_iterations"):
            assert gs[k] == cs[k], (block, k, gs[k], cs[k])
        for wi in (0, W // 2, W - 1):
            assert gpu.contact_pairs(wi) == cpu.contact_pairs(wi), (block, wi)
        gc, cc = gpu.contacts(W // 3), cpu.contacts(W // 3)       # the as-generated contacts of the block's last frame, bit for bit
        for f in ("point", "normal", "penetration", "friction", "restitution"):
            assert np.array_equal(gc.valid(f), cc.valid(f)), (block, f)
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS + ("transform", "inverse_inertia_tensor_world", "last_frame_acceleration"):
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    assert np.array_equal(gpu.download_colliders().transform, cpu.download_colliders().transform)
    assert gpu.checksum_energy()[0] == cpu.checksum_energy()[0]
    gpu.close()


@pytest.mark.parametrize("prec", [_abi.F64, _abi.F32], ids=["f64", "f32"])
@pytest.mark.parametrize("path", ["fused8", "split-lanes", "fused16", "multi"])
def test_random_worlds_with_materials_episodes_and_forces(path, prec, monkeypatch):
    """The same random worlds with everything the world handle adds on top of the reference's loop switched on together:
    per-pair surface materials, episode resets at staggered phases, and force / torque accumulators fed between calls."""
    from cubez_b200.api import BatchedWorld
    B, W, cap, env, flags = PATHS[path]
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    scene = scenes.with_materials(scenes.random_worlds(prec, n_worlds=W, bodies_per_world=B, seed=29, n_planes=2), seed=5)
    gpu = BatchedWorld.from_scene(scene, flags=flags, contacts_per_world=cap)
    cpu = OracleWorld.from_scene(scene)
    phase0 = (np.arange(W) * 11) % 70
    gpu.set_episodes(70, phase0)
    cpu.set_episodes(70, phase0)
    rng = np.random.default_rng(9)
    for block, n in enumerate((1, 30, 2, 45, 1, 60)):
        if block % 2 == 0:   # a push and a twist for one body in five
            sel = rng.uniform(0, 1, (W * B, 1)) < 0.2
            f, t = rng.uniform(-40, 40, (W * B, 3)) * sel, rng.uniform(-5, 5, (W * B, 3)) * sel
            gpu.add_forces(f, t)
            cpu.add_forces(f, t)
        gs, cs = gpu.step(scene.dt, n), cpu.step(scene.dt, n, n_threads=8)
        assert cs["status"] == 0 and cs["max_contacts"] <= cap
        for k in ("contacts", "pos_iterations", "vel_iterations"):
            assert gs[k] == cs[k], (block, k, gs[k], cs[k])
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS + ("transform", "inverse_inertia_tensor_world", "last_frame_acceleration"):
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    for wi in (0, W - 1):
        assert gpu.contact_pairs(wi) == cpu.contact_pairs(wi), wi
    gpu.close()


@pytest.mark.parametrize("islands", ["0", "1", "2"])
@pytest.mark.parametrize("n_bodies,seed,prec", [(300, 7, _abi.F64), (700, 8, _abi.F64), (300, 9, _abi.F32)], ids=["300-f64", "700-f64", "300-f32"])
def test_one_large_random_world_through_the_broadphase(n_bodies, seed, prec, islands, monkeypatch):
    """ONE large world of random cubes, spheres and collider-less bodies of different sizes, a third of them with a
    rotated and shifted collider Offset, three half-spaces: the sort-based broadphase (cell size from the largest bound),
    the plane pass and the large-world resolver (single CTA, or islands) against the oracle's O(n^2) loop — contact
    sequence every five frames, state bits at the end."""
    from cubez_b200.api import BatchedWorld
    monkeypatch.setenv("CUBEZ_RESOLVE_ISLANDS", islands)
    scene = scenes.random_worlds(prec, n_worlds=1, bodies_per_world=n_bodies, seed=seed, n_planes=3, extent=5.0, height=9.0)
    gpu = BatchedWorld.from_scene(scene, flags=_abi.WORLD_BROADPHASE, contacts_per_world=8 * n_bodies)
    cpu = OracleWorld.from_scene(scene)
    for s in range(0, 40, 5):
        gs, cs = gpu.step(scene.dt, 5), cpu.step(scene.dt, 5)
        assert (gs["contacts"], gs["pos_iterations"], gs["vel_iterations"]) == (cs["contacts"], cs["pos_iterations"], cs["vel_iterations"]), s
        assert gpu.contact_pairs(0) == cpu.contact_pairs(0), s
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS:
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    gpu.close()


@pytest.mark.parametrize("path", ["fused8", "split-lanes", "multi"])
def test_random_worlds_rl_loop_and_host_step(path, monkeypatch):
    """Random worlds through the two host-facing steps: the RL step (AddVelocity + AddRotation on every body, episode
    resets, observations back) and the host-resident step (state up, frame, state down), against the oracle doing the
    same edits on its own copy."""
    from cubez_b200.api import BatchedWorld, Context
    B, W, cap, env, flags = PATHS[path]
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    scene = scenes.random_worlds(_abi.F64, n_worlds=W, bodies_per_world=B, seed=57, n_planes=2)
    ctx = Context.get(0, "f64")
    gpu = BatchedWorld.from_scene(scene, flags=flags, contacts_per_world=cap)
    cpu = OracleWorld.from_scene(scene)
    phase0 = (np.arange(W) * 5) % 45
    gpu.set_episodes(45, phase0)
    cpu.set_episodes(45, phase0)
    nb = W * B
    rng = np.random.default_rng(2)
    av, ar = ctx.pinned_array((nb, 3)), ctx.pinned_array((nb, 3))
    obs = ctx.pinned_bodies(nb, fields=BatchedWorld.OBS_FIELDS)
    for it, n in enumerate((1, 2, 25, 1, 40, 3)):
        a = rng.uniform(-0.4, 0.4, (nb, 3)) * (rng.uniform(0, 1, (nb, 1)) < 0.15)
        r = rng.uniform(-0.6, 0.6, (nb, 3)) * (rng.uniform(0, 1, (nb, 1)) < 0.15)
        av[...] = a
        ar[...] = r
        gs = gpu.step_rl(av, ar, obs, scene.dt, n)
        d = cpu.download()
        d.velocity[...] = d.velocity + a.astype(d.velocity.dtype)
        d.rotation[...] = d.rotation + r.astype(d.rotation.dtype)
        cpu.upload_bodies(d)
        cs = cpu.step(scene.dt, n, n_threads=8)
        for k in ("contacts", "pos_iterations", "vel_iterations"):
            assert gs[k] == cs[k], (it, k, gs[k], cs[k])
        c = cpu.download()
        for f in BatchedWorld.OBS_FIELDS:
            assert np.array_equal(getattr(obs, f), getattr(c, f)), (it, f)
    # host-resident stepping from here on: the host copy is edited (a nudge), uploaded, stepped, downloaded
    host = gpu.download(out=ctx.pinned_bodies(nb))
    for it in range(4):
        nudge = rng.uniform(-0.05, 0.05, (nb, 3)) * (rng.uniform(0, 1, (nb, 1)) < 0.1)
        host.velocity[...] = host.velocity + nudge
        d = cpu.download()
        d.velocity[...] = d.velocity + nudge
        cpu.upload_bodies(d)
        gs = gpu.step_host(host, scene.dt, 1)
        cs = cpu.step(scene.dt, 1, n_threads=8)
        for k in ("contacts", "pos_iterations", "vel_iterations"):
            assert gs[k] == cs[k], ("host", it, k)
        c = cpu.download()
        for f in STATE_FIELDS + ("transform", "inverse_inertia_tensor_world", "last_frame_acceleration"):
            assert np.array_equal(getattr(host, f), getattr(c, f)), ("host", it, f)
    gpu.close()


@pytest.mark.parametrize("prec", [_abi.F64, _abi.F32], ids=["f64", "f32"])
@pytest.mark.parametrize("path", ["fused8", "split-lanes", "fused16", "multi"])
def test_random_explicit_schedule(path, prec, monkeypatch):
    """An explicit check list drawn at random — both orders of a pair, repeated entries (each repeat appends its contacts
    again, as calling CheckForCollisions twice does), a plane as the first or the second operand (colliders.go:116-125
    hands (plane, body) to the body's own check), plane against plane (no contact, :111-113)."""
    from cubez_b200.api import BatchedWorld
    B, W, cap, env, flags = PATHS[path]
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    scene = scenes.random_worlds(prec, n_worlds=W, bodies_per_world=B, seed=71, n_planes=2)
    rng = np.random.default_rng(B)
    one, two = [], []
    for _ in range(5 * B):
        kind = rng.uniform()
        a, b2 = (int(v) for v in rng.choice(B, 2, replace=False))
        pl = -1 - int(rng.integers(0, 2))
        if kind < 0.55:
            one.append(a); two.append(b2)
        elif kind < 0.75:
            one.append(a); two.append(pl)
        elif kind < 0.92:
            one.append(pl); two.append(a)
        else:
            one.append(-1); two.append(-2)
        if rng.uniform() < 0.15:            # the same check again, right behind
            one.append(one[-1]); two.append(two[-1])
    scene.schedule = _abi.SCHED_EXPLICIT
    scene.check_one, scene.check_two = np.asarray(one, dtype=np.int32), np.asarray(two, dtype=np.int32)
    cap = max(cap, 160)
    gpu = BatchedWorld.from_scene(scene, flags=flags, contacts_per_world=cap)
    cpu = OracleWorld.from_scene(scene)
    for block in range(5):
        gs, cs = gpu.step(scene.dt, 20), cpu.step(scene.dt, 20, n_threads=8)
        assert cs["status"] == 0
        for k in ("contacts", "pos_iterations", "vel_iterations"):
            assert gs[k] == cs[k], (block, k, gs[k], cs[k])
        for wi in (0, W - 1):
            assert gpu.contact_pairs(wi) == cpu.contact_pairs(wi), (block, wi)
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS:
        assert np.array_equal(getattr(g, f), getattr(c, f), equal_nan=True), f
    gpu.close()


@pytest.mark.parametrize("path", ["fused8", "split-lanes", "multi"])
def test_random_worlds_partial_uploads_mid_run(path, monkeypatch):
    """Edits in the middle of a run: a range of worlds is downloaded, changed on the host and uploaded again — primary
    state only with derived data recomputed on the device (derive = 1), or the whole record as it is (derive = 0) — while
    the other worlds go on untouched; the oracle gets the same edits."""
    from cubez_b200.api import BatchedWorld
    B, W, cap, env, flags = PATHS[path]
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    scene = scenes.random_worlds(_abi.F64, n_worlds=W, bodies_per_world=B, seed=83, n_planes=2)
    gpu = BatchedWorld.from_scene(scene, flags=flags, contacts_per_world=cap)
    cpu = OracleWorld.from_scene(scene)
    rng = np.random.default_rng(4)
    for it, (first, n, derive) in enumerate(((3, 10, True), (W - 7, 7, False), (0, W, True), (W // 2, 1, False))):
        gs, cs = gpu.step(scene.dt, 25), cpu.step(scene.dt, 25, n_threads=8)
        for k in ("contacts", "pos_iterations", "vel_iterations"):
            assert gs[k] == cs[k], (it, k)
        d = gpu.download(first, n)
        e = cpu.download(first, n)
        for f in STATE_FIELDS + ("transform", "inverse_inertia_tensor_world", "last_frame_acceleration"):
            assert np.array_equal(getattr(d, f), getattr(e, f)), (it, f)
        # the edit: lift and spin every body of the range, wake it
        d.position[:, 1] += rng.uniform(0.5, 1.5, d.position.shape[0]).astype(d.position.dtype)
        d.rotation[...] = rng.uniform(-1, 1, d.rotation.shape).astype(d.rotation.dtype)
        d.is_awake[:] = 1
        d.motion[:] = 0.6
        gpu.upload_bodies(d, first_world=first, derive=derive)
        cpu.upload_bodies(d, first_world=first, derive=derive)
    gs, cs = gpu.step(scene.dt, 30), cpu.step(scene.dt, 30, n_threads=8)
    for k in ("contacts", "pos_iterations", "vel_iterations"):
        assert gs[k] == cs[k], k
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS + ("transform", "inverse_inertia_tensor_world", "last_frame_acceleration"):
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    assert np.array_equal(gpu.download_colliders().transform, cpu.download_colliders().transform)
    gpu.close()
