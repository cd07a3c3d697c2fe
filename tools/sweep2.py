import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cubez_b200 import scenes
from cubez_b200.api import BatchedWorld
W = 65536
sc = scenes.batched_cubedrop(n_worlds=W)
ph = (np.arange(W) % 600).astype(np.int32)
for LS in ("3", "2", "1", "0"):
    os.environ["CUBEZ_FUSED_LOCKSTEP"] = LS
    gpu = BatchedWorld.from_scene(sc, contacts_per_world=64)
    gpu.set_episodes(600, ph)
    gpu.step(sc.dt, 600)
    st = gpu.step(sc.dt, 60)
    print(f"LOCKSTEP={LS}: {st['device_ms']/60:.3f} ms/frame -> {W*60/st['device_ms']/1e3:.2f} M ws/s", flush=True)
    gpu.close()
