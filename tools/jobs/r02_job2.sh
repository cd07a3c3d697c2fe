#!/bin/bash
# round-2 GPU job 2: the adjacency-list large-world resolver: parity (broadphase, cfg3, reference dumps) and timing against the round-1 loop
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_broadphase.py tests/test_gpu_cfg3.py tests/test_gpu_vs_reference_dump.py tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/r02_job2_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_job2_tests.log
for mode in 3 2; do
  echo "== CUBEZ_RESOLVE_MODE=$mode" >> gpurun_out/r02_cfg3_modes.log
  CUBEZ_RESOLVE_MODE=$mode timeout 300 python tools/cfg3_probe.py --load tests/golden/pile4096_f100.npz --steps 3 >> gpurun_out/r02_cfg3_modes.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resolve -c 1 -f -o gpurun_out/r02_resolve_settled_v3 python tools/cfg3_probe.py --load tests/golden/pile4096_f100.npz --steps 1 > gpurun_out/r02_resolve_ncu_v3.log 2>&1
tail -6 gpurun_out/r02_job2_tests.log; cat gpurun_out/r02_cfg3_modes.log
