import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cubez_b200 import scenes, _abi
from cubez_b200.api import BatchedWorld
side = int(os.environ.get("PILE_SIDE", 16))
blocks = int(os.environ.get("PILE_BLOCKS", 8))
sc = scenes.pile(side=side)
gpu = BatchedWorld.from_scene(sc, flags=_abi.WORLD_BROADPHASE, contacts_per_world=8 * side ** 3)
tot = 0
for blk in range(blocks):
    t = time.perf_counter()
    st = gpu.step(sc.dt, 10)
    el = time.perf_counter() - t
    tot += 10
    print(f"frames {tot-10}-{tot}: {el/10*1e3:.2f} ms/frame, contacts/frame {st['contacts']/10:.0f}, pos it/frame {st['pos_iterations']/10:.0f}, vel it/frame {st['vel_iterations']/10:.0f}, max contacts {st['max_contacts']}", flush=True)
    if el > 40: break
print("checksum", hex(gpu.checksum_energy()[0]))
