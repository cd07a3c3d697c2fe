import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cubez_b200.api import Context
ctx = Context.get(0, "f64")
n = 1 << 24
ms, pairs, sms = C.c_float(), C.c_int64(), C.c_float()
ctx.check(ctx.lib.cz_bench_broadphase(ctx.h, n, 7, 0.05, 1, 2, C.byref(ms), C.byref(pairs), C.byref(sms)))
print(ms.value, sms.value, pairs.value)
