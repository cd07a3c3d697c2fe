#!/bin/bash
# ncu captures behind profiles/ncu_counters.json (DRAM bytes and FP64 instruction counts of the code as built)
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum
# one frame (A | B | C) of the stationary 65 536-world population as ONE lane (CUBEZ_STEP_LANES=1: same instructions and bytes as the
# two-lane step, three launches per frame to pick): the profile script runs 600 + 1 + 1 frames, 3 launches each
CUBEZ_STEP_LANES=1 PROF_WORLDS=65536 PROF_FRAMES=1 timeout 900 ncu --metrics $M --clock-control none --print-units base --csv --page raw -k regex:k_world_fused --launch-skip 1803 --launch-count 3 --log-file gpurun_out/ncu_fused.csv python tools/profile_fused.py > gpurun_out/ncu_fused.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --print-units base --csv --page raw -k regex:k_integrate --launch-skip 3 --launch-count 1 --log-file gpurun_out/ncu_k1.csv python -c "
import sys; sys.path.insert(0,'.')
from cubez_b200.api import Context
print(Context.get(0,'f64').bench_integrate(1<<24, warmup=3, steps=2))" > gpurun_out/ncu_k1.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --print-units base --csv --page raw -k regex:'k_bp_|scan' --launch-skip 0 --launch-count 40 --log-file gpurun_out/ncu_k2_all.csv python -c "
import sys, ctypes as C; sys.path.insert(0,'.')
from cubez_b200.api import Context
ctx = Context.get(0,'f64'); ms, p, s = C.c_float(), C.c_int64(), C.c_float()
ctx.check(ctx.lib.cz_bench_broadphase(ctx.h, 1<<24, 7, 0.05, 0, 1, C.byref(ms), C.byref(p), C.byref(s)))" > gpurun_out/ncu_k2.log 2>&1
