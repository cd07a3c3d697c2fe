"""Resident split-mode step as 1 / 2 / 3 / 4 independent slices on their own streams (CUBEZ_STEP_LANES)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cubez_b200 import scenes
from cubez_b200.api import BatchedWorld
for W in [int(a) for a in sys.argv[1:]] or (65536, 32768):
    sc = scenes.batched_cubedrop(n_worlds=W)
    ph = (np.arange(W) % 600).astype(np.int32)
    for lanes in (1, 2, 3, 4):
        os.environ["CUBEZ_STEP_LANES"] = str(lanes)
        gpu = BatchedWorld.from_scene(sc, contacts_per_world=64)
        gpu.set_episodes(600, ph)
        gpu.step(sc.dt, 600)
        st = gpu.step(sc.dt, 200)
        cs = gpu.checksum_energy()[0]
        gpu.close()
        print(f"W={W} lanes={lanes}: {st['device_ms']/200*1e3:.1f} us per frame = {W*200/st['device_ms']/1e3:.2f} M world-steps/s, checksum {cs:#x}", flush=True)
