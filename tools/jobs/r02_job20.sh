#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/lanes_probe.py > gpurun_out/r02_lanes_probe.log 2>&1; cat gpurun_out/r02_lanes_probe.log
timeout 600 python tools/rl_async_probe.py 1 2 3 2>&1 | tail -4
