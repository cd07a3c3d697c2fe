#!/usr/bin/env python
"""Compare a dump printed by one of go/harness/*_headless.go — run with a REAL Go toolchain against the unmodified
reference — with the CPU oracle (and, with --gpu, the CUDA path), frame by frame and bit for bit.  This is how anyone
with Go closes the last link of the parity chain (the Go compiler itself; everything else is already pinned by the
mechanically translated reference, tests/golden/ref/).

    go run go/harness/cubedrop_headless.go 600            > d.txt ; python tools/compare_go_dump.py d.txt
    go run go/harness/cubedrop_headless.go 600 256 0      > d.txt ; python tools/compare_go_dump.py d.txt --worlds 256
    go run go/harness/ballistic_headless.go 600           > d.txt ; python tools/compare_go_dump.py d.txt
    go run go/harness/pile_headless.go 60 16              > d.txt ; python tools/compare_go_dump.py d.txt [--gpu]
    go run go/harness/integrate_bench_headless.go 16 65536 > d.txt ; python tools/compare_go_dump.py d.txt
    go run go/harness/random_headless.go 120 8 8 11 2     > d.txt ; python tools/compare_go_dump.py d.txt [--gpu]

The header line of the dump names the scene; --worlds / --first-world / --bullets / --second-fire repeat the harness
arguments that the header does not carry.  A committed dump of the translated reference can be compared the same way
(python tools/compare_go_dump.py tests/golden/ref/ballistic_600.txt), and two dumps with `diff`."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refdump  # noqa: E402
from cubez_b200 import _abi, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dump")
    ap.add_argument("--gpu", action="store_true", help="compare the CUDA path instead of the CPU oracle")
    ap.add_argument("--worlds", type=int, default=0)
    ap.add_argument("--first-world", type=int, default=0)
    ap.add_argument("--bullets", type=int, default=64)
    ap.add_argument("--second-fire", type=int, default=0)
    a = ap.parse_args()
    text = open(a.dump).read()
    header, frames, _ = refdump.parse(text)
    name, bodies = header.get("scene"), int(header.get("bodies", 0))
    if name == "cubedrop":
        scene = scenes.batched_cubedrop(n_worlds=a.worlds, first_world=a.first_world) if a.worlds else scenes.cubedrop(second_fire_step=a.second_fire)
    elif name == "ballistic":
        scene = scenes.ballistic(n_bullets=bodies - 2)
    elif name == "pile":
        scene = scenes.pile(side=round(bodies ** (1 / 3)))
    elif name == "free_bodies":
        scene = scenes.free_bodies(n=bodies)
    elif name == "random_worlds":
        worlds = int(header.get("worlds", 1))
        big = int(header.get("big", 0)) > 0        # the harness's sixth argument: one large world, spread out
        scene = scenes.random_worlds(n_worlds=worlds, bodies_per_world=bodies // worlds, seed=int(header.get("seed", 11)), n_planes=int(header.get("planes", 2)),
                                     **({"extent": 5.0, "height": 9.0} if big else {}))
        if int(header.get("materials", 0)) > 0:
            scene = scenes.with_materials(scene, seed=int(header["materials"]))
    else:
        raise SystemExit(f"unknown scene in the dump header: {header}")
    if a.gpu:
        from cubez_b200.api import BatchedWorld
        world = BatchedWorld.from_scene(scene, flags=_abi.WORLD_BROADPHASE if name == "pile" or (name == "random_worlds" and scene.n_worlds == 1 and scene.bodies_per_world >= 100) else 0)
    else:
        from oracle_lib import OracleWorld
        world = OracleWorld.from_scene(scene)
    lines = refdump.run_dump(world, scene, len(frames))
    diff = refdump.first_difference(text, lines)
    print(f"{len(frames)} frames of '{name}' ({bodies} bodies) compared with the {'CUDA path' if a.gpu else 'CPU oracle'}: "
          + ("IDENTICAL (contact counts, pair sequences, contact geometry, every body's state bits)" if diff is None else "DIFFERENT — " + diff))
    sys.exit(0 if diff is None else 1)


if __name__ == "__main__":
    main()
