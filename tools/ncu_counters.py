"""Turn ncu CSV captures (--csv --page raw) of THIS build into profiles/ncu_counters.json, the file bench.py reads the
measured DRAM traffic and FP64 instruction counts from (keyed by a hash of cubez_b200/csrc: a stale capture reads as null).
  python tools/ncu_counters.py --fused fused.csv --worlds 65536 [--k1 k1.csv] [--k2 k2.csv]
The captures are made by tools/jobs/ncu_counters.sh on the GPU box."""
import argparse, csv, io, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import csrc_sha  # noqa: E402


def rows(path):
    text = open(path, errors="replace").read()
    start = text.find('"ID"')
    r = list(csv.reader(io.StringIO(text[start:])))
    hdr = r[0]
    return [dict(zip(hdr, x)) for x in r[2:] if len(x) == len(hdr)]


def num(v):
    try:
        return float(str(v).replace(",", ""))
    except ValueError:
        return 0.0


def dram(rs):
    tot = 0.0
    for r in rs:
        for k, mult in (("dram__bytes_read.sum", 1), ("dram__bytes_write.sum", 1)):
            tot += num(r.get(k, 0))
    return tot


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--fused"); ap.add_argument("--worlds", type=int, default=65536)
    ap.add_argument("--k1"); ap.add_argument("--k2")
    a = ap.parse_args()
    out = {"csrc_sha": csrc_sha(), "made_by": "tools/ncu_counters.py from ncu --page raw --csv captures (tools/jobs/ncu_counters.sh)"}
    path = os.path.join(ROOT, "profiles", "ncu_counters.json")
    if os.path.exists(path):
        old = json.load(open(path))
        if old.get("csrc_sha") == out["csrc_sha"]:
            out.update(old)
    if a.fused:
        rs = [r for r in rows(a.fused) if "k_world_fused" in r.get("Kernel Name", "")]
        unit = {}
        ops = 0.0
        for r in rs:
            for k in ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"):
                ops += num(r.get(k, 0))
        out["fused_frame_launches"] = len(rs)
        out["fused_frame_dram_bytes"] = dram(rs)   # NOTE: ncu reports bytes in the unit of the column; captured with --print-units base
        out["fused_fp64_ops_per_world_step"] = ops / a.worlds
        out["fused_frame_ms_under_ncu"] = sum(num(r.get("gpu__time_duration.sum", 0)) for r in rs) / 1e6
    if a.k1:
        rs = [r for r in rows(a.k1) if "k_integrate" in r.get("Kernel Name", "")]
        out["k1_dram_bytes"] = dram(rs[-1:])
    if a.k2:
        rs = rows(a.k2)
        names = [r.get("Kernel Name", "") for r in rs]
        first = next(i for i, n in enumerate(names) if "k_bp_count" in n or "k_bp_bin" in n)
        last = next(i for i, n in enumerate(names) if "k_bp_sweep" in n)
        rs = rs[first:last + 1]      # one frame of the candidate pipeline (without the sphere generator and the radix sort timed beside it)
        out["k2_dram_bytes"] = dram(rs)   # kernels only: the cudaMemsetAsync of the cell table (4 B per cell) is not a kernel
        out["k2_kernels"] = [r.get("Kernel Name", "")[:40] for r in rs]
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))
