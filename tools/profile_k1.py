import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cubez_b200.api import Context
for prec in ("f64", "f32"):
    ms, ck = Context.get(0, prec).bench_integrate(1 << 24, warmup=2, steps=3)
    print(prec, ms)
