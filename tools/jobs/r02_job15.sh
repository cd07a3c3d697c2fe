#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rl_step.py tests/test_gpu_worlds.py -m gpu -x -q > gpurun_out/r02_job15_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_job15_tests.log
tail -4 gpurun_out/r02_job15_tests.log
{
echo "--- batched copies"; python tools/e2e_probe.py 6 8 12 16 2>&1 | tail -4
echo "--- plain copies"; CUBEZ_COPY_BATCH=0 python tools/e2e_probe.py 6 8 2>&1 | tail -2
echo "--- batched, trace"; CUBEZ_HOST_TRACE=1 python tools/e2e_probe.py 8 2>&1 | tail -10
echo "--- rl"; python tools/rl_probe.py 2>&1 | tail -8
} > gpurun_out/r02_e2e_batch.log 2>&1
cat gpurun_out/r02_e2e_batch.log
