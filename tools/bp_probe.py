"""K2 microbench probe: 16 Mi unit spheres at 5 % fill, per-stage trace, both sorts, grid sweeps."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CUBEZ_BP_TRACE"] = "1"
from cubez_b200.api import Context
ctx = Context(0, "f64")
n = 1 << 24
def run(label):
    ms, pairs, sms = C.c_float(), C.c_int64(), C.c_float()
    ctx.check(ctx.lib.cz_bench_broadphase(ctx.h, n, 7, 0.05, 2, 3, C.byref(ms), C.byref(pairs), C.byref(sms)))
    alg = 148 * n + 8 * pairs.value
    print(f"{label} n={n}: {ms.value:.3f} ms/frame, sort alone {sms.value:.3f} ms, pairs {pairs.value}, {alg/ms.value/1e6:.0f} GB/s = {alg/ms.value/1e6/6560:.2f} of measured HBM", flush=True)
os.environ["CUBEZ_BP_SORT"] = "count"
for cpb in sys.argv[1:] or ("0.5", "1", "2", "4"):
    os.environ["CUBEZ_BP_CELLS_PER_BODY"] = cpb
    run(f"count cells/body {cpb}")
os.environ["CUBEZ_BP_SORT"] = "radix"
run("radix")
