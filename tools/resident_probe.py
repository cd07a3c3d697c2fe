import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cubez_b200 import scenes
from cubez_b200.api import BatchedWorld, Context
W = 65536
sc = scenes.batched_cubedrop(n_worlds=W)
gpu = BatchedWorld.from_scene(sc, contacts_per_world=int(os.environ.get("CAP", 64)))
gpu.set_episodes(600, (np.arange(W) % 600).astype(np.int32))
gpu.step(sc.dt, 600)
st = gpu.step(sc.dt, 20)
print(f"resident: {st['device_ms']/20:.3f} ms/frame max_contacts {st['max_contacts']}", {k: v for k, v in os.environ.items() if k.startswith('CUBEZ')})
