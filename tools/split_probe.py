import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cubez_b200 import scenes
from cubez_b200.api import BatchedWorld
W = 65536
sc = scenes.batched_cubedrop(n_worlds=W)
ph = (np.arange(W) % 600).astype(np.int32)
ref = None
os.environ["CUBEZ_FUSED_SPLIT"] = "1"
for mb, th, cap in (("2", "128", 64), ("3", "128", 48), ("3", "64", 48), ("4", "64", 48), ("4", "64", 64), ("4", "32", 48), ("3", "32", 48)):
    os.environ["CUBEZ_FUSED_SPLIT_MINB"] = mb; os.environ["CUBEZ_FUSED_THREADS"] = th
    gpu = BatchedWorld.from_scene(sc, contacts_per_world=cap)
    gpu.set_episodes(600, ph)
    gpu.step(sc.dt, 600)
    st = gpu.step(sc.dt, 60)
    ck = gpu.checksum_energy()[0]; ref = ref or ck
    print(f"SPLIT_MINB={mb} THREADS={th} cap={cap}: {st['device_ms']/60:.3f} ms/frame -> {W*60/st['device_ms']/1e3:.2f} M ws/s, checksum same {ck == ref}", flush=True)
    gpu.close()
