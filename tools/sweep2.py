import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cubez_b200 import scenes
from cubez_b200.api import BatchedWorld
W = 65536
sc = scenes.batched_cubedrop(n_worlds=W)
ph = (np.arange(W) % 600).astype(np.int32)
for MB, cap in (("2", 64), ("2", 48), ("3", 48), ("4", 48), ("3", 47)):
    os.environ["CUBEZ_FUSED_MINB"] = MB
    gpu = BatchedWorld.from_scene(sc, contacts_per_world=cap)
    gpu.set_episodes(600, ph)
    gpu.step(sc.dt, 600)
    st = gpu.step(sc.dt, 60)
    print(f"MINB={MB} cap={cap}: {st['device_ms']/60:.3f} ms/frame -> {W*60/st['device_ms']/1e3:.2f} M ws/s, max contacts {st['max_contacts']}", flush=True)
    gpu.close()
