"""GPU: force / torque accumulator input (cz_world_add_forces).  The reference keeps
forceAccum / torqueAccum, reads them in Integrate (rigidbody.go:219-223) and clears them (:206), but has no writer; the
oracle restates exactly those lines, so the device path is compared with it bit for bit: fused and multi-kernel paths,
f64 and f32, accumulation over several calls, and sleeping bodies keeping their accumulators until they wake."""
import numpy as np
import pytest

from cubez_b200 import _abi, scenes
from golden_cases import STATE_FIELDS
from oracle_lib import OracleWorld

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("flags", [_abi.WORLD_FUSED, _abi.WORLD_NO_FUSED])
def test_forces_and_torques_match_oracle(prec, flags):
    from cubez_b200.api import BatchedWorld
    p = _abi.precision(prec)
    scene = scenes.batched_cubedrop(p, n_worlds=40)
    gpu = BatchedWorld.from_scene(scene, flags=flags)
    cpu = OracleWorld.from_scene(scene)
    rng = np.random.default_rng(5)
    n = scene.n_bodies
    for frame in range(160):
        if frame % 3 != 2:                       # some frames without input: the accumulators were cleared
            f = rng.uniform(-30, 30, (n, 3)).astype(p.dtype)
            t = rng.uniform(-2, 2, (n, 3)).astype(p.dtype)
            if frame % 7 == 0:                   # accumulate twice before the step; force only
                gpu.add_forces(f, None); cpu.add_forces(f, None)
            gpu.add_forces(f, t); cpu.add_forces(f, t)
        gs, cs = gpu.step(scene.dt, 1), cpu.step(scene.dt, 1)
        for k in ("contacts", "pos_iterations", "vel_iterations"):
            assert gs[k] == cs[k], (frame, k, gs[k], cs[k])
    g, c = gpu.download(), cpu.download()
    for f_ in STATE_FIELDS + ("last_frame_acceleration", "transform"):
        assert np.array_equal(getattr(g, f_), getattr(c, f_)), f_
    gpu.close()


def test_sleeping_body_keeps_its_accumulators_until_it_wakes():
    """Integrate returns before ClearAccumulators when IsAwake is false (rigidbody.go:214-216): a force added to a
    sleeping body is applied on the first frame after something wakes it."""
    from cubez_b200.api import BatchedWorld
    scene = scenes.cubedrop()
    gpu, cpu = BatchedWorld.from_scene(scene), OracleWorld.from_scene(scene)
    gpu.step(scene.dt, 450); cpu.step(scene.dt, 450)
    assert not gpu.download().is_awake.any()
    f = np.zeros((8, 3)); f[:, 1] = 500.0
    gpu.add_forces(f, None); cpu.add_forces(f, None)
    gpu.step(scene.dt, 5); cpu.step(scene.dt, 5)
    g, c = gpu.download(), cpu.download()
    assert not g.is_awake.any()                                  # nothing woke them: the force is still pending
    for f_ in STATE_FIELDS:
        assert np.array_equal(getattr(g, f_), getattr(c, f_)), f_
    # a part-world upload that wakes body 0: its pending force now acts
    b = gpu.download()
    b.is_awake[0] = 1
    b.motion[0] = 2.0                                            # (SetAwake(true) would give it 0.6: enough not to doze off at once)
    gpu.upload_bodies(b); cpu.upload_bodies(b)
    gpu.step(scene.dt, 1); cpu.step(scene.dt, 1)
    assert gpu.download().velocity[0, 1] > 0.5                   # (500 / 8 - 9.78) / 60 m/s upwards: the pending force acted, once
    gpu.step(scene.dt, 29); cpu.step(scene.dt, 29)
    g, c = gpu.download(), cpu.download()
    for f_ in STATE_FIELDS:
        assert np.array_equal(getattr(g, f_), getattr(c, f_)), f_
    gpu.close()
