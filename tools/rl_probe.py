import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cubez_b200 import scenes
from cubez_b200.api import BatchedWorld, Context
W = 65536
sc = scenes.batched_cubedrop(n_worlds=W)
ctx = Context.get(0, "f64")
nb = W * 8
act = ctx.pinned_array((nb, 3)); act[...] = np.random.default_rng(1).uniform(-1e-3, 1e-3, (nb, 3))
obs = ctx.pinned_bodies(nb, fields=BatchedWorld.OBS_FIELDS)
gpu = BatchedWorld.from_scene(sc, contacts_per_world=64)
gpu.set_episodes(600, (np.arange(W) % 600).astype(np.int32))
gpu.step(sc.dt, 600)
for chunks in [int(a) for a in sys.argv[1:]] or (1, 2, 3, 4, 6, 8):
    os.environ["CUBEZ_RL_CHUNKS"] = str(chunks)
    gpu.step_rl(act, None, obs, sc.dt, 1)
    t = time.perf_counter(); dev = 0
    for _ in range(10):
        dev += gpu.step_rl(act, None, obs, sc.dt, 1)["device_ms"]
    el = (time.perf_counter() - t) / 10
    print(f"rl chunks={chunks}: wall {el*1e3:.2f} ms/frame, device events {dev/10:.2f} ms", flush=True)
