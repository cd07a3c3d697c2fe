#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rl_step.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/rl_async_probe.py > gpurun_out/r02_rl_async_chunks.log 2>&1; cat gpurun_out/r02_rl_async_chunks.log
timeout 600 python tools/strong_probe.py 16384 2>&1 | head -2
