"""one configuration of the K2 microbench (for ncu): python tools/bp_probe1.py [n_log2] [warmup] [steps]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cubez_b200.api import Context
ctx = Context.get(0, "f64")
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1
k = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ms, pairs, sms = C.c_float(), C.c_int64(), C.c_float()
ctx.check(ctx.lib.cz_bench_broadphase(ctx.h, n, 7, 0.05, w, k, C.byref(ms), C.byref(pairs), C.byref(sms)))
print(ms.value, sms.value, pairs.value)
