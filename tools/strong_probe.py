"""BASELINE cfg4's strong-scaling share of one GPU: W = 65 536 / N worlds, 600 frames from t = 0 in one call.
Which fused-kernel configuration is fastest at each W?  (env knobs of czf::plan)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cubez_b200 import scenes
from cubez_b200.api import BatchedWorld
VARIANTS = [{}, {"CUBEZ_FUSED_SPLIT": "1"}, {"CUBEZ_FUSED_MINB": "3"}, {"CUBEZ_FUSED_G": "16"}, {"CUBEZ_FUSED_LOCKSTEP": "0"}, {"CUBEZ_FUSED_LOCKSTEP": "1"},
            {"CUBEZ_FUSED_SPLIT": "1", "CUBEZ_FUSED_SPLIT_MINB": "2", "CUBEZ_FUSED_PHASE_A_MINB": "2"}]
KEYS = sorted({k for v in VARIANTS for k in v})
for W in [int(a) for a in sys.argv[1:]] or (8192, 16384, 32768):
    sc = scenes.batched_cubedrop(n_worlds=W)
    for var in VARIANTS:
        for k in KEYS: os.environ.pop(k, None)
        os.environ.update(var)
        gpu = BatchedWorld.from_scene(sc, contacts_per_world=64)
        st = gpu.step(sc.dt, 600)
        cs = gpu.checksum_energy()[0]
        gpu.close()
        print(f"W={W} {var or 'default'}: {st['device_ms']:.1f} ms for 600 frames = {W*600/st['device_ms']/1e3:.2f} M world-steps/s, checksum {cs:#x}", flush=True)
