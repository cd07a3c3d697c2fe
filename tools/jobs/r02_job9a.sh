#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/forces_probe.py > gpurun_out/r02_forces_probe.log 2>&1
cat gpurun_out/r02_forces_probe.log
timeout 900 python -m pytest tests/test_gpu_forces.py tests/test_gpu_broadphase.py tests/test_gpu_rl_step.py -m gpu -q > gpurun_out/r02_job9a_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_job9a_tests.log
tail -6 gpurun_out/r02_job9a_tests.log
