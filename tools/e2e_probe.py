import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cubez_b200 import scenes
from cubez_b200.api import BatchedWorld, Context
W = 65536
sc = scenes.batched_cubedrop(n_worlds=W)
ctx = Context.get(0, "f64")
for chunks in [int(a) for a in sys.argv[1:]] or (4, 8, 12, 16):
    os.environ["CUBEZ_HOST_CHUNKS"] = str(chunks)
    gpu = BatchedWorld.from_scene(sc, contacts_per_world=64)
    gpu.set_episodes(600, (np.arange(W) % 600).astype(np.int32))
    gpu.step(sc.dt, 600)
    host = gpu.download(out=ctx.pinned_bodies(W * 8))
    gpu.step_host(host, sc.dt, 1)
    t = time.perf_counter(); dev = 0
    for _ in range(10):
        dev += gpu.step_host(host, sc.dt, 1)["device_ms"]
    el = (time.perf_counter() - t) / 10
    st = gpu.step(sc.dt, 10)
    print(f"chunks={chunks}: wall {el*1e3:.2f} ms/frame, device events {dev/10:.2f} ms; resident step {st['device_ms']/10:.2f} ms/frame", flush=True)
    gpu.close()
