#!/bin/bash
# round-2 GPU job 10: compute-sanitizer racecheck / synccheck over every kernel family (tools/sanitize_scenes.py), memcheck over the GPU tests file by file
mkdir -p gpurun_out
for tool in racecheck synccheck; do
  for which in fused8 fused16 fused32 split resolve32 resolve256 islands; do
    echo "=== $tool $which" >> gpurun_out/r02_sanitizer_$tool.log
    timeout 420 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_scenes.py $which >> gpurun_out/r02_sanitizer_$tool.log 2>&1
    echo "rc $?" >> gpurun_out/r02_sanitizer_$tool.log
  done
done
for f in test_gpu_kernels test_gpu_islands test_gpu_forces test_gpu_math_kat test_gpu_run_abi; do
  echo "=== memcheck $f" >> gpurun_out/r02_sanitizer_memcheck.log
  timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/$f.py -m gpu -x -q >> gpurun_out/r02_sanitizer_memcheck.log 2>&1
  echo "rc $?" >> gpurun_out/r02_sanitizer_memcheck.log
done
grep -E "===|ERROR SUMMARY|rc |passed|failed" gpurun_out/r02_sanitizer_racecheck.log gpurun_out/r02_sanitizer_synccheck.log gpurun_out/r02_sanitizer_memcheck.log | tail -70
