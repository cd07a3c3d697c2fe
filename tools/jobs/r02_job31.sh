#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fuzz_worlds.py tests/test_gpu_forces.py tests/test_gpu_rl_step.py -m gpu -q > gpurun_out/r02_fuzz_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_fuzz_tests.log
tail -12 gpurun_out/r02_fuzz_tests.log
timeout 300 python tools/fuzz_isolate.py 2>&1 | grep -v "None, final state equal True" | head
