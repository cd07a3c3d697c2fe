"""ncu target: stationary cfg4 population, one profiled k_world_fused launch of N frames."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cubez_b200 import scenes
from cubez_b200.api import BatchedWorld
W = int(os.environ.get("PROF_WORLDS", 16384))
frames = int(os.environ.get("PROF_FRAMES", 8))
sc = scenes.batched_cubedrop(n_worlds=W)
gpu = BatchedWorld.from_scene(sc, contacts_per_world=64)
gpu.set_episodes(600, (np.arange(W) % 600).astype(np.int32))
gpu.step(sc.dt, 600)          # launch 0: pre-roll
gpu.step(sc.dt, frames)       # launch 1: warm
st = gpu.step(sc.dt, frames)  # launch 2: profiled
print("frames", frames, "worlds", W, "ms", st["device_ms"], "M ws/s", W * frames / st["device_ms"] / 1e3)
