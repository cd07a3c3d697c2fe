"""Which feature makes the combined fuzz scene differ from the oracle?  (diagnostic)"""
import os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cubez_b200 import _abi, scenes
from cubez_b200.api import BatchedWorld
from oracle_lib import OracleWorld
B, W, cap = 8, 96, 64
for mats, eps, forces, act in itertools.product((0, 1), repeat=4):
    scene = scenes.random_worlds(_abi.F64, n_worlds=W, bodies_per_world=B, seed=29, n_planes=2)
    if not act: scene.active_from[:] = 0
    if mats: scene = scenes.with_materials(scene, seed=5)
    gpu = BatchedWorld.from_scene(scene, flags=_abi.WORLD_NO_FUSED, contacts_per_world=cap)
    cpu = OracleWorld.from_scene(scene)
    if eps:
        phase0 = (np.arange(W) * 11) % 70
        gpu.set_episodes(70, phase0); cpu.set_episodes(70, phase0)
    rng = np.random.default_rng(9)
    bad = None
    for block, n in enumerate((1, 30, 2, 45, 1, 60)):
        if forces and block % 2 == 0:
            sel = rng.uniform(0, 1, (W * B, 1)) < 0.2
            f, t = rng.uniform(-40, 40, (W * B, 3)) * sel, rng.uniform(-5, 5, (W * B, 3)) * sel
            gpu.add_forces(f, t); cpu.add_forces(f, t)
        gs, cs = gpu.step(scene.dt, n), cpu.step(scene.dt, n, n_threads=8)
        if any(gs[k] != cs[k] for k in ("contacts", "pos_iterations", "vel_iterations")) and bad is None:
            bad = (block, gs["contacts"], cs["contacts"])
    g, c = gpu.download(), cpu.download()
    same = all(np.array_equal(getattr(g, f), getattr(c, f)) for f in ("position", "velocity", "is_awake"))
    print(f"materials={mats} episodes={eps} forces={forces} activation={act}: first counter mismatch {bad}, final state equal {same}", flush=True)
    gpu.close()
