"""CPU, world_size 2 over gloo: the multi-GPU plumbing of bench.py (world sharding by rank,
max-over-ranks timing, checksum / energy / counter all-reduce) with the oracle standing in for
the device step.  No data-path collective exists: worlds never exchange state."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
from cubez_b200 import scenes
from cubez_b200.sharding import shard_range, reduce_run
from oracle_lib import OracleWorld
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
first, n = shard_range(48, rank, 2)
sc = scenes.batched_cubedrop(n_worlds=n, first_world=first)
w = OracleWorld.from_scene(sc)
st = w.step(sc.dt, 120)
cks, en = w.checksum_energy()
out = reduce_run(cks, en, {"contacts": st["contacts"], "vel_iterations": st["vel_iterations"]}, elapsed_ms=10.0 + rank, device=torch.device("cpu"))
if rank == 0:
    print("RESULT", out["checksum"], repr(out["energy"]), out["counters"]["contacts"], out["counters"]["vel_iterations"], out["max_ms"])
dist.destroy_process_group()
"""


def test_two_rank_sharding_and_reduce():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cubez_b200 import scenes
    from cubez_b200.sharding import shard_range
    from oracle_lib import OracleWorld
    assert shard_range(48, 0, 2) == (0, 24) and shard_range(48, 1, 2) == (24, 24)
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]
    port = 29500 + os.getpid() % 2000
    code = WORKER % {"root": ROOT, "port": port}
    procs = [subprocess.Popen([sys.executable, "-c", code, str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT")][0].split()
    sc = scenes.batched_cubedrop(n_worlds=48)
    w = OracleWorld.from_scene(sc)
    st = w.step(sc.dt, 120)
    cks, en = w.checksum_energy()
    assert int(line[1]) == cks
    assert abs(float(line[2]) - en) <= 1e-9 * abs(en)
    assert int(line[3]) == st["contacts"] and int(line[4]) == st["vel_iterations"]
    assert float(line[5]) == 11.0          # max over ranks
