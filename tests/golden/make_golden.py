"""Generates the golden fixtures under tests/golden/ from the CPU oracle.

These .npz fixtures are ORACLE-DERIVED regression pins (per-world counters, iteration counts, final state; also the
float32 and materials cases, which the reference cannot produce): they keep the oracle and the CUDA path from drifting
silently.  Independent evidence of matching the reference is elsewhere: tests/golden/ref/*.txt are dumps printed by the
reference's own Go sources (translated mechanically, oracle/go2cpp.py) and both the oracle and the CUDA path reproduce
them bit for bit.   Usage:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from cubez_b200 import _abi, scenes  # noqa: E402
from hostemu_lib import pair_hash  # noqa: E402
from oracle_lib import OracleWorld  # noqa: E402

CASES = {
    "cubedrop_f64": (lambda: scenes.cubedrop(_abi.F64), 200),
    "cubedrop_f32": (lambda: scenes.cubedrop(_abi.F32), 200),
    "cubedrop_staggered_f64": (lambda: scenes.cubedrop(_abi.F64, second_fire_step=120), 260),
    "ballistic16_f64": (lambda: scenes.ballistic(_abi.F64, n_bullets=16), 300),
    "batched8_f64": (lambda: scenes.batched_cubedrop(_abi.F64, n_worlds=8), 150),
    "pile27_f64": (lambda: scenes.pile(_abi.F64, side=3), 150),
    # per-pair surface materials (cz_world_set_materials): asymmetric 3x3 tables, ids per body, slippery ground
    "cubedrop_materials_f64": (lambda: scenes.with_materials(scenes.cubedrop(_abi.F64), seed=3), 220),
    "batched6_materials_f64": (lambda: scenes.with_materials(scenes.batched_cubedrop(_abi.F64, n_worlds=6)), 200),
    "pile27_materials_f32": (lambda: scenes.with_materials(scenes.pile(_abi.F32, side=3), seed=11), 150),
}


def run_case(make, n_steps):
    sc = make()
    w = OracleWorld.from_scene(sc)
    counts = np.zeros((n_steps, sc.n_worlds), dtype=np.int32)
    pos = np.zeros_like(counts)
    vel = np.zeros_like(counts)
    ph = np.zeros((n_steps, sc.n_worlds), dtype=np.uint64)
    for s in range(n_steps):
        w.step(sc.dt, 1)
        c, p, v = w.last_counts()
        counts[s], pos[s], vel[s] = c, p, v
        for k in range(sc.n_worlds):
            ph[s, k] = pair_hash(w.contact_pairs(k))
    b = w.download()
    cks, en = w.checksum_energy()
    out = dict(counts=counts, pos_iters=pos, vel_iters=vel, pair_hash=ph, checksum=np.uint64(cks), energy=np.float64(en))
    for f in ("position", "orientation", "velocity", "rotation", "motion", "is_awake"):
        out[f] = getattr(b, f)
    return out


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, (make, n) in CASES.items():
        if only and name not in only:
            continue
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_case(make, n))
        print("wrote", name)
