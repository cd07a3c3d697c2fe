#!/bin/bash
# round-2 GPU job 6: contact islands (parity + timing), ncu --set full of the three fused phase launches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_islands.py tests/test_gpu_cfg3.py tests/test_gpu_broadphase.py tests/test_gpu_vs_reference_dump.py -m gpu -x -q > gpurun_out/r02_job6_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_job6_tests.log
for isl in 1 0; do
  echo "== CUBEZ_RESOLVE_ISLANDS=$isl" >> gpurun_out/r02_cfg3_islands.log
  CUBEZ_RESOLVE_ISLANDS=$isl timeout 300 python tools/cfg3_probe.py --frames 50 --block 5 >> gpurun_out/r02_cfg3_islands.log 2>&1
done
PROF_WORLDS=65536 PROF_FRAMES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_world_fused --launch-skip 1803 --launch-count 3 -f -o gpurun_out/r02_fused_phases python tools/profile_fused.py > gpurun_out/r02_fused_phases_ncu.log 2>&1
tail -5 gpurun_out/r02_job6_tests.log; cat gpurun_out/r02_cfg3_islands.log
