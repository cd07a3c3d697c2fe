import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cubez_b200.api import Context
ctx = Context(0, "f64")
n = 1 << 24
for scale in ("1", "1.3", "1.6", "2"):
    os.environ["CUBEZ_BP_CELL_SCALE"] = scale
    ms, pairs, sms = C.c_float(), C.c_int64(), C.c_float()
    ctx.check(ctx.lib.cz_bench_broadphase(ctx.h, n, 7, 0.05, 2, 3, C.byref(ms), C.byref(pairs), C.byref(sms)))
    alg = 148 * n + 8 * pairs.value
    print(f"scale {scale} n={n}: {ms.value:.3f} ms/frame, sort alone {sms.value:.3f} ms, pairs {pairs.value}, {alg/ms.value/1e6:.0f} GB/s = {alg/ms.value/1e6/6560:.2f} of measured HBM", flush=True)
