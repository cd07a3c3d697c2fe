"""Pipelined RL step (cz_world_step_rl_async, two steps in flight, float32 observations): chunks of the step that is
enqueued behind another one (CUBEZ_RL_ASYNC_CHUNKS)."""
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
acts = [ctx.pinned_array((nb, 3)) for _ in range(2)]
for a in acts: a[...] = np.random.default_rng(1).uniform(-1e-3, 1e-3, (nb, 3))
obs32 = [{k: ctx.pinned_array((nb, c), dtype=np.float32) for k, c in (("position", 3), ("orientation", 4), ("velocity", 3), ("rotation", 3))} for _ in range(2)]
gpu = BatchedWorld.from_scene(sc, contacts_per_world=64)
gpu.set_episodes(600, (np.arange(W) % 600).astype(np.int32))
gpu.step(sc.dt, 600)
def run(n):
    tickets = []
    for f in range(n):
        if f >= 2: gpu.rl_wait(tickets[f - 2], stats=False)
        tickets.append(gpu.step_rl_async(acts[f & 1], None, None, obs32[f & 1], sc.dt, 1))
    for t in tickets[-2:]: gpu.rl_wait(t, stats=False)
for chunks in [int(a) for a in sys.argv[1:]] or (1, 2, 4):
    os.environ["CUBEZ_RL_ASYNC_CHUNKS"] = str(chunks)
    run(4)
    t = time.perf_counter(); run(40); el = (time.perf_counter() - t) / 40
    print(f"async chunks={chunks}: {el*1e3:.3f} ms per step = {W/el/1e6:.2f} M world-steps/s", flush=True)
st = gpu.step(sc.dt, 20); print("resident", st["device_ms"] / 20)
