#!/bin/bash
# round-2 GPU job 3: cz_run_* ABI, resolver with precomputed velocity response + redux arg-max (parity + timing), K2 stage trace
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_run_abi.py tests/test_gpu_broadphase.py tests/test_gpu_cfg3.py tests/test_gpu_vs_reference_dump.py tests/test_gpu_kernels.py tests/test_gpu_object_api.py -m gpu -x -q > gpurun_out/r02_job3_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_job3_tests.log
rm -f gpurun_out/r02_cfg3_modes_b.log
for pre in 0 1; do
  echo "== mode 3, CUBEZ_RESOLVE_NO_PRE=$pre" >> gpurun_out/r02_cfg3_modes_b.log
  CUBEZ_RESOLVE_NO_PRE=$pre timeout 300 python tools/cfg3_probe.py --load tests/golden/pile4096_f100.npz --steps 3 >> gpurun_out/r02_cfg3_modes_b.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resolve -c 1 -f -o gpurun_out/r02_resolve_settled_v4 python tools/cfg3_probe.py --load tests/golden/pile4096_f100.npz --steps 1 > gpurun_out/r02_resolve_ncu_v4.log 2>&1
timeout 300 python tools/bp_probe.py 1 > gpurun_out/r02_k2_trace_before.log 2>&1
tail -6 gpurun_out/r02_job3_tests.log; cat gpurun_out/r02_cfg3_modes_b.log; tail -12 gpurun_out/r02_k2_trace_before.log
