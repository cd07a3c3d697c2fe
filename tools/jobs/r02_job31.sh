#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fuzz_worlds.py -m gpu -q > gpurun_out/r02_fuzz_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_fuzz_tests.log
tail -30 gpurun_out/r02_fuzz_tests.log
