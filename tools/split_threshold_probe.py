"""Persistent kernel vs one launch per phase (split mode) at small world counts: where is the crossover?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cubez_b200 import scenes
from cubez_b200.api import BatchedWorld
for W in [int(a) for a in sys.argv[1:]] or (512, 1024, 2048, 4096, 8192):
    sc = scenes.batched_cubedrop(n_worlds=W)
    ph = (np.arange(W) % 600).astype(np.int32)
    res = []
    for split in ("0", "1"):
        os.environ["CUBEZ_FUSED_SPLIT"] = split
        gpu = BatchedWorld.from_scene(sc, contacts_per_world=64)
        gpu.set_episodes(600, ph)
        gpu.step(sc.dt, 600)
        st = gpu.step(sc.dt, 120)
        res.append((st["device_ms"] / 120, gpu.checksum_energy()[0]))
        gpu.close()
    print(f"W={W}: persistent {res[0][0]*1e3:.1f} us/frame, split {res[1][0]*1e3:.1f} us/frame, same checksum {res[0][1] == res[1][1]}", flush=True)
