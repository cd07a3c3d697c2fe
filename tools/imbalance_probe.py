"""How much of the lockstep loops is idle?  quad = the 4 worlds of a warp; cost = max over the quad."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cubez_b200 import scenes
from cubez_b200.api import BatchedWorld
W = 65536
sc = scenes.batched_cubedrop(n_worlds=W)
gpu = BatchedWorld.from_scene(sc, contacts_per_world=64)
gpu.set_episodes(600, (np.arange(W) % 600).astype(np.int32))
gpu.step(sc.dt, 600)
gpu.step(sc.dt, 7)
_, p0, v0 = [np.asarray(a) for a in gpu.last_counts()]
gpu.step(sc.dt, 1)
c1, p1, v1 = [np.asarray(a) for a in gpu.last_counts()]
def eff(actual, order):
    a = actual[order].reshape(-1, 4)
    return actual.sum() / (4 * a.max(axis=1).sum())
ident = np.arange(W)
for name, prev, act, sh in (("pos", p0, p1, 1), ("vel", v0, v1, 2)):
    pred = np.argsort(-(np.minimum(prev >> sh, 63)), kind="stable")
    perfect = np.argsort(-act, kind="stable")
    print(f"{name}: mean {act.mean():.2f} max {act.max()} nonzero {np.mean(act>0):.2f} | lockstep efficiency identity {eff(act, ident):.2f}, predicted order {eff(act, pred):.2f}, oracle order {eff(act, perfect):.2f}")
    print("   corr(prev, act) =", np.corrcoef(prev, act)[0, 1])
