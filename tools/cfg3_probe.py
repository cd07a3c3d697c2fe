"""cfg3 (4 096-body pile through the broadphase) on the GPU: per-frame time, contacts and iteration counts;
optionally saves / restarts from a snapshot of the piled state (body + collider state through the ABI's
download / upload).   python tools/cfg3_probe.py [--frames N] [--save PATH --at F] [--load PATH --steps K]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cubez_b200 import scenes, _abi
from cubez_b200.api import BatchedWorld

SNAP_B = ("position", "orientation", "velocity", "rotation", "motion", "is_awake", "transform", "inverse_inertia_tensor_world", "last_frame_acceleration")


def save_snapshot(path, gpu, frame):
    b, c = gpu.download(), gpu.download_colliders()
    np.savez_compressed(path, frame=np.int64(frame), collider_transform=c.transform, **{f: getattr(b, f) for f in SNAP_B})


def load_snapshot(path, gpu, scene):
    z = np.load(path)
    b, c = scene.bodies, scene.colliders
    for f in SNAP_B:
        getattr(b, f)[...] = z[f]
    c.transform[...] = z["collider_transform"]
    gpu.upload_bodies(b, derive=False)
    gpu.upload_colliders(c, derive=False)
    gpu.set_step_index(int(z["frame"]))
    return int(z["frame"])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=80)
    ap.add_argument("--block", type=int, default=5)
    ap.add_argument("--save"); ap.add_argument("--at", type=int, default=-1)
    ap.add_argument("--load"); ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--side", type=int, default=16)
    a = ap.parse_args()
    sc = scenes.pile(side=a.side)
    gpu = BatchedWorld.from_scene(sc, flags=_abi.WORLD_BROADPHASE)
    if a.load:
        f0 = load_snapshot(a.load, gpu, sc)
        for k in range(a.steps):
            t = time.perf_counter(); st = gpu.step(sc.dt, 1); el = time.perf_counter() - t
            print(f"frame {f0 + k}: {el * 1e3:.2f} ms wall, {st['device_ms']:.2f} ms device, contacts {st['contacts']}, pos it {st['pos_iterations']}, vel it {st['vel_iterations']}", flush=True)
        print("checksum", hex(gpu.checksum_energy()[0]))
        sys.exit(0)
    tot, t_all = 0, time.perf_counter()
    while tot < a.frames:
        n = min(a.block, a.frames - tot)
        if a.save and tot < a.at <= tot + n:
            n = a.at - tot
        t = time.perf_counter(); st = gpu.step(sc.dt, n); el = time.perf_counter() - t
        tot += n
        it = st['pos_iterations'] + st['vel_iterations']
        print(f"frames {tot - n}-{tot}: {el / n * 1e3:.2f} ms/frame, contacts/frame {st['contacts'] / n:.0f}, pos it/frame {st['pos_iterations'] / n:.0f}, vel it/frame {st['vel_iterations'] / n:.0f}, us/iteration {el * 1e6 / max(it, 1):.2f}", flush=True)
        if a.save and tot == a.at:
            save_snapshot(a.save, gpu, tot)
            print("snapshot saved at frame", tot, flush=True)
    print(f"total {time.perf_counter() - t_all:.1f} s for {tot} frames; checksum", hex(gpu.checksum_energy()[0]))
