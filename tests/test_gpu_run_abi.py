"""GPU: the in-library multi-GPU entry points (cz_run_*): worlds sharded over devices inside one process, no
per-step exchange, one grouped ncclAllReduce of {checksum, energy, counters, time} at the end.  The reduced checksum
must equal the single-handle checksum and the CPU oracle's whatever the shard count (worlds are independent and the
checksum is an integer sum).  With one visible GPU the shards share device 0 (host reduce); with two or more the real
single-process NCCL communicator is exercised."""
import ctypes as C

import numpy as np
import pytest

from cubez_b200 import scenes
from oracle_lib import OracleWorld

pytestmark = pytest.mark.gpu


def _device_count():
    from cubez_b200 import _abi
    import ctypes
    rt = ctypes.CDLL("libcudart.so")
    n = ctypes.c_int()
    assert rt.cudaGetDeviceCount(ctypes.byref(n)) == 0
    return n.value


def _single_handle(scene, frames):
    from cubez_b200.api import BatchedWorld
    w = BatchedWorld.from_scene(scene, contacts_per_world=64)
    st = w.step(scene.dt, frames)
    out = (w.checksum_energy(), st["contacts"], st["pos_iterations"], st["vel_iterations"])
    w.close()
    return out


@pytest.mark.parametrize("shards", [1, 2, 3])
def test_run_shards_on_one_device_equal_single_handle_and_oracle(shards):
    from cubez_b200.api import Run
    scene = scenes.batched_cubedrop(n_worlds=96)
    frames = 150
    (cks, energy), contacts, pos, vel = _single_handle(scene, frames)
    run = Run(scene, [0] * shards, contacts_per_world=64)
    assert [run.shard(k)[1:] for k in range(shards)] == [(96 * k // shards, 96 * (k + 1) // shards - 96 * k // shards) for k in range(shards)]
    run.step(scene.dt, 100)
    run.step(scene.dt, 50)
    t = run.finish()
    run.close()
    assert t["checksum"] == cks and t["world_steps"] == 96 * frames
    assert (t["contacts"], t["pos_iterations"], t["vel_iterations"]) == (contacts, pos, vel)
    assert abs(t["energy"] - energy) <= 1e-9 * abs(energy)
    assert t["n_shards"] == shards and t["used_nccl"] == 0 and t["max_device_ms"] > 0
    cpu = OracleWorld.from_scene(scene)
    cpu.step(scene.dt, frames, n_threads=8)
    assert cpu.checksum_energy()[0] == cks


def test_run_over_all_devices_with_nccl():
    n = _device_count()
    if n < 2:
        pytest.skip("needs two or more GPUs (gpurun --gpus 2)")
    from cubez_b200.api import Run
    scene = scenes.batched_cubedrop(n_worlds=64 * n)
    frames = 120
    (cks, energy), contacts, pos, vel = _single_handle(scene, frames)
    run = Run(scene, list(range(n)), contacts_per_world=64)
    run.step(scene.dt, frames)
    t = run.finish()
    assert t["used_nccl"] == 1 and t["n_shards"] == n
    assert t["checksum"] == cks and (t["contacts"], t["pos_iterations"], t["vel_iterations"]) == (contacts, pos, vel)
    assert abs(t["energy"] - energy) <= 1e-9 * abs(energy)
    # a second round on the same run: counters and timer were reset by finish
    run.step(scene.dt, 10)
    t2 = run.finish()
    assert t2["world_steps"] == 64 * n * 10
    run.close()


def test_run_rejects_bad_arguments():
    from cubez_b200._abi import CubezError
    from cubez_b200.api import Run
    scene = scenes.batched_cubedrop(n_worlds=2)
    with pytest.raises(CubezError):
        Run(scene, [0, 0, 0])           # fewer worlds than shards
    with pytest.raises(CubezError):
        Run(scene, [99])                # no such device
