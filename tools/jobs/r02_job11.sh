#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_broadphase.py tests/test_gpu_cfg3.py tests/test_gpu_islands.py tests/test_gpu_worlds.py::test_nonfinite_body_count tests/test_gpu_forces.py -m gpu -x -q > gpurun_out/r02_job11_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_job11_tests.log
tail -4 gpurun_out/r02_job11_tests.log
timeout 300 python tools/cfg3_probe.py --load tests/golden/pile4096_f100.npz --steps 3 2>&1 | tee gpurun_out/r02_cfg3_settled_v5.log
