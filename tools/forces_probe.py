import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cubez_b200 import scenes
from cubez_b200.api import BatchedWorld
from oracle_lib import OracleWorld
F = ("position", "orientation", "velocity", "rotation", "motion", "is_awake")
def diff(tag, gpu, cpu):
    g, c = gpu.download(), cpu.download()
    bad = [f for f in F if not np.array_equal(getattr(g, f), getattr(c, f))]
    print(tag, "DIFF" if bad else "same", bad, "gpu y", g.position[:, 1].tolist() if bad else "", "cpu y", c.position[:, 1].tolist() if bad else "", flush=True)
for trial in range(3):
    for chunk in (450, 50):
        scene = scenes.cubedrop()
        gpu, cpu = BatchedWorld.from_scene(scene), OracleWorld.from_scene(scene)
        for s in range(0, 450, chunk):
            gpu.step(scene.dt, chunk); cpu.step(scene.dt, chunk)
        diff(f"trial {trial} chunk {chunk}: after 450", gpu, cpu)
        gpu.step(scene.dt, 5); cpu.step(scene.dt, 5)
        diff("   +5 without forces", gpu, cpu)
        f = np.zeros((8, 3)); f[:, 1] = 500.0
        gpu.add_forces(f, None); cpu.add_forces(f, None)
        gpu.step(scene.dt, 5); cpu.step(scene.dt, 5)
        diff("   +5 with pending forces", gpu, cpu)
        for k in range(5):
            gpu.step(scene.dt, 1); cpu.step(scene.dt, 1)
        diff("   +5 single frames", gpu, cpu)
        gpu.close()
