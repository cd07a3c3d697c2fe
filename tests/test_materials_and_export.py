"""per-pair surface materials (replacing the hard-wired `c.Friction = 0.9` /
`c.Restitution = 0.1` test constants of colliders.go:199-202 and five more sites) and the renderer-side
float32 export (examples/cubedrop.go:35-37).  CPU part: the oracle's semantics; GPU part: the CUDA paths
(multi-kernel, fused persistent, fused split-phase, sort-based broadphase, RL step) against the oracle."""
import os

import numpy as np
import pytest

from cubez_b200 import _abi, scenes
from golden_cases import STATE_FIELDS
from oracle_lib import OracleWorld


# ---- CPU: the oracle's definition --------------------------------------------------------------
def test_oracle_uniform_table_equals_reference_constants():
    """A table that holds 0.9 / 0.1 everywhere is the reference's behaviour, bit for bit."""
    sc = scenes.batched_cubedrop(n_worlds=3)
    a = OracleWorld.from_scene(sc)
    sc2 = scenes.with_materials(scenes.batched_cubedrop(n_worlds=3))
    sc2.materials["friction"][:] = 0.9
    sc2.materials["restitution"][:] = 0.1
    b = OracleWorld.from_scene(sc2)
    sa, sb = a.step(sc.dt, 200), b.step(sc.dt, 200)
    assert (sa["contacts"], sa["pos_iterations"], sa["vel_iterations"]) == (sb["contacts"], sb["pos_iterations"], sb["vel_iterations"])
    assert a.checksum_energy()[0] == b.checksum_energy()[0]


def test_oracle_contacts_carry_the_table_entry_and_materials_change_the_motion():
    sc = scenes.with_materials(scenes.batched_cubedrop(n_worlds=4))
    m = sc.materials
    w = OracleWorld.from_scene(sc)
    ref = OracleWorld.from_scene(scenes.batched_cubedrop(n_worlds=4))
    seen = 0
    for s in range(0, 180, 6):
        w.step(sc.dt, 6)
        for k in range(4):
            c = w.contacts(k)
            ids = m["body_material"][k * 8:(k + 1) * 8]
            for i in range(c.count):
                b0, b1 = int(c.body0[i]), int(c.body1[i])
                if b1 >= 0:
                    continue          # pair contacts: Bodies may be (two, one) (colliders.go:636-640); the table is asymmetric
                seen += 1             # plane contacts: check (collider b0, plane 0)
                assert c.friction[i] == m["friction"][ids[b0], m["plane_material"][0]]
                assert c.restitution[i] == m["restitution"][ids[b0], m["plane_material"][0]]
    assert seen > 100
    ref.step(sc.dt, 180)
    assert w.checksum_energy()[0] != ref.checksum_energy()[0]


def test_oracle_frictionless_one_body_contact_is_the_reference_panic():
    """Friction == 0 selects calculateFrictionlessImpulse, whose second-body block dereferences a nil body for a
    plane contact (contact.go:512-523): reported as a status, never silently computed."""
    sc = scenes.with_materials(scenes.cubedrop())
    sc.materials["friction"][:] = 0.0
    w = OracleWorld.from_scene(sc)
    st = w.step(sc.dt, 120)
    assert st["status"] == _abi.CZ_ERR_NIL_BODY


# ---- GPU ------------------------------------------------------------------------------------------
def _gpu_world(scene, flags=0, env=None, **kw):
    from cubez_b200.api import BatchedWorld
    env = env or {}
    for k, v in env.items():
        os.environ[k] = v
    try:
        return BatchedWorld.from_scene(scene, flags=flags, **kw)
    finally:
        for k in env:
            os.environ.pop(k, None)


GPU_PATHS = {
    "multi": (_abi.WORLD_NO_FUSED, {}),
    "fused8": (_abi.WORLD_FUSED, {"CUBEZ_FUSED_G": "8", "CUBEZ_FUSED_SPLIT": "0"}),
    "fused8_split": (_abi.WORLD_FUSED, {"CUBEZ_FUSED_G": "8", "CUBEZ_FUSED_SPLIT": "1"}),
    "fused16": (_abi.WORLD_FUSED, {"CUBEZ_FUSED_G": "16", "CUBEZ_FUSED_SPLIT": "0"}),
    "fused32_split": (_abi.WORLD_FUSED, {"CUBEZ_FUSED_G": "32", "CUBEZ_FUSED_SPLIT": "1"}),
}


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("path", sorted(GPU_PATHS))
def test_gpu_materials_match_oracle(path, prec):
    """40 worlds with three materials: per-call counters, contact records (incl. Friction / Restitution) and the
    final state are the oracle's, bit for bit."""
    P = _abi.precision(prec)
    sc = scenes.with_materials(scenes.batched_cubedrop(P, n_worlds=40))
    flags, env = GPU_PATHS[path]
    gpu, cpu = _gpu_world(sc, flags, env), OracleWorld.from_scene(sc)
    for s in range(0, 240, 20):
        gs, cs = gpu.step(sc.dt, 20), cpu.step(sc.dt, 20, n_threads=4)
        for k in ("contacts", "pos_iterations", "vel_iterations"):
            assert gs[k] == cs[k], (s, k, gs[k], cs[k])
        for wi in (0, 17, 39):
            gc, cc = gpu.contacts(wi), cpu.contacts(wi)
            assert gc.count == cc.count
            for f in ("body0", "body1", "friction", "restitution", "point", "normal", "penetration"):
                assert np.array_equal(gc.valid(f), cc.valid(f)), (s, wi, f)
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS:
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    assert gpu.checksum_energy()[0] == cpu.checksum_energy()[0]
    gpu.close()


@pytest.mark.gpu
def test_gpu_materials_reset_and_uniform_table():
    """friction=None restores the constants; a uniform 0.9 / 0.1 table equals them."""
    base = scenes.batched_cubedrop(n_worlds=16)
    w0 = _gpu_world(base)
    w0.step(base.dt, 150)
    want = w0.checksum_energy()[0]
    w0.close()
    sc = scenes.with_materials(scenes.batched_cubedrop(n_worlds=16))
    w1 = _gpu_world(sc)
    w1.set_materials(None, None)
    w1.step(sc.dt, 150)
    assert w1.checksum_energy()[0] == want
    w1.close()
    sc.materials["friction"][:] = 0.9
    sc.materials["restitution"][:] = 0.1
    w2 = _gpu_world(sc)
    w2.step(sc.dt, 150)
    assert w2.checksum_energy()[0] == want
    w2.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["multi", "fused8"])
def test_gpu_frictionless_one_body_contact_is_an_error(path):
    sc = scenes.with_materials(scenes.cubedrop())
    sc.materials["friction"][:] = 0.0
    flags, env = GPU_PATHS[path]
    w = _gpu_world(sc, flags, env)
    with pytest.raises(_abi.CubezError) as ei:
        w.step(sc.dt, 120)
    assert ei.value.code == _abi.CZ_ERR_NIL_BODY
    w.close()


@pytest.mark.gpu
def test_gpu_materials_through_the_sort_based_broadphase():
    """One 64-body pile of cubes and spheres with materials, stepped through K2: the contact records are recovered
    from the canonical sort key and must carry the oracle's Friction / Restitution."""
    sc = scenes.with_materials(scenes.pile(_abi.F64, side=4), seed=5)
    gpu, cpu = _gpu_world(sc, _abi.WORLD_BROADPHASE | _abi.WORLD_NO_FUSED, contacts_per_world=2048), OracleWorld.from_scene(sc)
    for s in range(0, 150, 10):
        gs, cs = gpu.step(sc.dt, 10), cpu.step(sc.dt, 10)
        for k in ("contacts", "pos_iterations", "vel_iterations"):
            assert gs[k] == cs[k], (s, k)
        gc, cc = gpu.contacts(0), cpu.contacts(0)
        for f in ("body0", "body1", "friction", "restitution", "penetration"):
            assert np.array_equal(gc.valid(f), cc.valid(f)), (s, f)
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS:
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    gpu.close()


@pytest.mark.gpu
def test_gpu_materials_on_the_explicit_ballistic_schedule():
    """cfg2's explicit check list (examples/ballistic.go:47-97; 1 058 checks with 16 bullets, planes and colliders on
    either side of a check): the table is indexed by the operands as the schedule names them."""
    sc = scenes.with_materials(scenes.ballistic(_abi.F64, n_bullets=16), seed=9)
    gpu, cpu = _gpu_world(sc, _abi.WORLD_NO_FUSED), OracleWorld.from_scene(sc)
    for s in range(0, 300, 25):
        gs, cs = gpu.step(sc.dt, 25), cpu.step(sc.dt, 25)
        for k in ("contacts", "pos_iterations", "vel_iterations"):
            assert gs[k] == cs[k], (s, k)
        gc, cc = gpu.contacts(0), cpu.contacts(0)
        for f in ("body0", "body1", "friction", "restitution", "penetration"):
            assert np.array_equal(gc.valid(f), cc.valid(f)), (s, f)
    g, c = gpu.download(), cpu.download()
    for f in STATE_FIELDS:
        assert np.array_equal(getattr(g, f), getattr(c, f)), f
    gpu.close()


@pytest.mark.gpu
def test_gpu_rl_step_with_materials_and_episodes():
    sc = scenes.with_materials(scenes.batched_cubedrop(n_worlds=300))
    gpu, cpu = _gpu_world(sc, contacts_per_world=64), OracleWorld.from_scene(sc)
    phase0 = (np.arange(300) * 11) % 90
    gpu.set_episodes(90, phase0); cpu.set_episodes(90, phase0)
    tot_g = tot_c = 0
    for s in range(30):
        gs = gpu.step_rl(None, None, None, sc.dt, 4)
        cs = cpu.step(sc.dt, 4, n_threads=8)
        tot_g += gs["vel_iterations"]; tot_c += cs["vel_iterations"]
    assert tot_g == tot_c
    assert gpu.checksum_energy()[0] == cpu.checksum_energy()[0]
    gpu.close()


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_gpu_export_gl_is_the_float32_cast_of_the_state(prec):
    """SetGlVector3 / SetGlQuat (examples/exampleapp.go:146-159): dst = float32(src) per component."""
    P = _abi.precision(prec)
    sc = scenes.batched_cubedrop(P, n_worlds=33)
    w = _gpu_world(sc)
    w.step(sc.dt, 97)
    b = w.download()
    loc, rot, mdl = w.export_gl(model=True)
    assert loc.dtype == np.float32 and loc.shape == (33 * 8, 3) and rot.shape == (33 * 8, 4)
    assert np.array_equal(loc, b.position.astype(np.float32))
    assert np.array_equal(rot, b.orientation.astype(np.float32))
    t = b.transform.reshape(-1, 4, 3).astype(np.float32)          # column-major 3x4: [column][row]
    want = np.zeros((33 * 8, 4, 4), dtype=np.float32)
    want[:, :, :3] = t
    want[:, 3, 3] = 1.0
    assert np.array_equal(mdl.reshape(-1, 4, 4), want)
    loc2, rot2 = w.export_gl(first_world=5, n_worlds=3)
    assert np.array_equal(loc2, loc[40:64]) and np.array_equal(rot2, rot[40:64])
    w.close()
