"""Small scenes for compute-sanitizer (racecheck / synccheck / memcheck): every kernel family once, a few frames each.
   compute-sanitizer --tool racecheck python tools/sanitize_scenes.py [which ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cubez_b200 import _abi, scenes
from cubez_b200.api import BatchedWorld

def run(name, scene, frames, flags=0, env=None, **kw):
    for k, v in (env or {}).items():
        os.environ[k] = v
    w = BatchedWorld.from_scene(scene, flags=flags, **kw)
    st = w.step(scene.dt, frames)
    print(name, "frames", frames, "contacts", st["contacts"], "pos", st["pos_iterations"], "vel", st["vel_iterations"], "checksum", hex(w.checksum_energy()[0]), flush=True)
    w.close()
    for k in (env or {}):
        os.environ.pop(k, None)

CASES = {
    "fused8": lambda: run("fused persistent G=8", scenes.batched_cubedrop(n_worlds=12), 110, _abi.WORLD_FUSED, {"CUBEZ_FUSED_G": "8"}),
    "fused16": lambda: run("fused persistent G=16", scenes.batched_cubedrop(n_worlds=6), 110, _abi.WORLD_FUSED, {"CUBEZ_FUSED_G": "16"}),
    "fused32": lambda: run("fused persistent G=32", scenes.ballistic(n_bullets=6), 100, _abi.WORLD_FUSED, {"CUBEZ_FUSED_G": "32"}),
    "split": lambda: run("fused split phases", scenes.batched_cubedrop(n_worlds=40), 105, _abi.WORLD_FUSED, {"CUBEZ_FUSED_SPLIT": "1", "CUBEZ_FUSED_G": "8"}, contacts_per_world=64),
    "lanes": lambda: run("fused split phases in 3 lanes", scenes.batched_cubedrop(n_worlds=48), 105, _abi.WORLD_FUSED,
                         {"CUBEZ_FUSED_SPLIT": "1", "CUBEZ_FUSED_G": "8", "CUBEZ_STEP_LANES": "3", "CUBEZ_STEP_LANE_MIN": "8"}, contacts_per_world=64),
    "resolve32": lambda: run("k_resolve<32> + k_narrow", scenes.cubedrop(), 110, _abi.WORLD_NO_FUSED),
    "resolve256": lambda: run("k_resolve<256> adjacency loop + broadphase", scenes.pile(side=4), 60, _abi.WORLD_BROADPHASE, {"CUBEZ_RESOLVE_ISLANDS": "0"}),
    "islands": lambda: run("k_resolve_islands", scenes.archipelago(piles=3, side=2), 70, _abi.WORLD_BROADPHASE, {"CUBEZ_RESOLVE_ISLANDS": "2"}),
}
if __name__ == "__main__":
    for which in (sys.argv[1:] or list(CASES)):
        CASES[which]()
