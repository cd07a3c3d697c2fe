#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_worlds.py tests/test_gpu_rl_step.py -m gpu -x -q 2>&1 | tail -3
for tool in racecheck synccheck memcheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_scenes.py lanes > gpurun_out/r02_sanitizer_lanes_$tool.log 2>&1
  tail -3 gpurun_out/r02_sanitizer_lanes_$tool.log
done
