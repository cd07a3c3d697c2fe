"""chunked compute without any transfer: cz_world_step_rl with no actions and no observation"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cubez_b200 import scenes
from cubez_b200.api import BatchedWorld, Context
W = 65536
sc = scenes.batched_cubedrop(n_worlds=W)
ctx = Context.get(0, "f64")
gpu = BatchedWorld.from_scene(sc, contacts_per_world=64)
gpu.set_episodes(600, (np.arange(W) % 600).astype(np.int32))
gpu.step(sc.dt, 600)
st = gpu.step(sc.dt, 10)
print(f"resident: {st['device_ms']/10:.2f} ms/frame")
for chunks in (1, 2, 4, 8):
    os.environ["CUBEZ_RL_CHUNKS"] = str(chunks)
    gpu.step_rl(None, None, None, sc.dt, 1)
    dev = 0
    for _ in range(10):
        dev += gpu.step_rl(None, None, None, sc.dt, 1)["device_ms"]
    print(f"compute only, chunks={chunks}: device events {dev/10:.2f} ms", flush=True)
