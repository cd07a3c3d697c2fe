"""BASELINE config 4 at full size on the CPU oracle: 65 536 perturbed cubedrop-8 worlds, 600 frames from t = 0
(SURVEY section 8d).  Writes tests/golden/cfg4_checksum.json: the FNV checksum summed over all worlds (mod 2^64), the
energy and the counters.  bench.py's strong-scaling arm must reproduce the checksum for every GPU count (worlds are
sharded 65 536 / N per rank), which ties the multi-GPU result of every world to the CPU restatement.
~1 minute on 8 host threads.   Usage: python tests/golden/make_cfg4_checksum.py [worlds]"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from cubez_b200 import scenes  # noqa: E402
from oracle_lib import OracleWorld  # noqa: E402

if __name__ == "__main__":
    W = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    frames, chunk = 600, 8192
    tot = {"checksum": 0, "energy": 0.0, "contacts": 0, "pos_iterations": 0, "vel_iterations": 0}
    t0 = time.time()
    for first in range(0, W, chunk):
        n = min(chunk, W - first)
        sc = scenes.batched_cubedrop(n_worlds=n, first_world=first)
        w = OracleWorld.from_scene(sc)
        st = w.step(sc.dt, frames, n_threads=os.cpu_count() or 1)
        cks, en = w.checksum_energy()
        tot["checksum"] = (tot["checksum"] + cks) & 0xFFFFFFFFFFFFFFFF
        tot["energy"] += en
        for k in ("contacts", "pos_iterations", "vel_iterations"):
            tot[k] += st[k]
        w.close()
        print(first + n, "worlds,", round(time.time() - t0, 1), "s", flush=True)
    out = {"worlds": W, "frames": frames, "checksum": hex(tot["checksum"]), "energy": tot["energy"], "contacts": tot["contacts"],
           "pos_iterations": tot["pos_iterations"], "vel_iterations": tot["vel_iterations"],
           "made_by": "tests/golden/make_cfg4_checksum.py (CPU oracle)"}
    name = "cfg4_checksum.json" if W == 65536 else f"cfg4_checksum_{W}.json"
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(out, f, indent=1)
    print(out)
