"""Golden fixture of the RL-style step (cz_world_step_rl), generated from the CPU
oracle driven as a host loop over the reference API would drive it: AddVelocity / AddRotation on
every body (rigidbody.go:195-202), then the frames.  Actions are a closed-form function of
(call, body, component) built from splitmix64, so the fixture needs no stored inputs.
Usage:  python tests/golden/make_golden_rl.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from cubez_b200 import _abi, scenes  # noqa: E402
from cubez_b200.hostmath import splitmix64_draws, uniform  # noqa: E402

N_WORLDS, CALLS, FRAMES_PER_CALL = 12, 40, 3
OBS = ("position", "orientation", "velocity", "rotation", "motion", "is_awake")


def make_scene(prec=_abi.F64):
    return scenes.batched_cubedrop(prec, n_worlds=N_WORLDS)


def actions(call: int, nb: int, dtype):
    """(add_velocity, add_rotation) of call `call`; None = no call (every third call has no rotation,
    every fifth no velocity).  One body in four is pushed."""
    u = splitmix64_draws(np.uint64(777 + call) + np.arange(nb, dtype=np.uint64), 7).reshape(nb, 7)
    push = (u[:, 6] < 0.25)[:, None]
    av = (uniform(u[:, 0:3], -0.4, 0.4) * push).astype(dtype)
    ar = (uniform(u[:, 3:6], -0.6, 0.6) * push).astype(dtype)
    return (None if call % 5 == 4 else av), (None if call % 3 == 2 else ar)


def run_oracle():
    from oracle_lib import OracleWorld
    sc = make_scene()
    w = OracleWorld.from_scene(sc)
    nb = N_WORLDS * sc.bodies_per_world
    counts = np.zeros((CALLS, 3), dtype=np.int64)
    for call in range(CALLS):
        av, ar = actions(call, nb, np.float64)
        d = w.download()
        if av is not None:
            d.velocity[...] = d.velocity + av
        if ar is not None:
            d.rotation[...] = d.rotation + ar
        w.upload_bodies(d)
        st = w.step(sc.dt, FRAMES_PER_CALL)
        counts[call] = (st["contacts"], st["pos_iterations"], st["vel_iterations"])
    b = w.download()
    out = dict(counts=counts, checksum=np.uint64(w.checksum_energy()[0]))
    for f in OBS:
        out[f] = getattr(b, f)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "rl12_f64.npz"), **run_oracle())
    print("wrote rl12_f64")
