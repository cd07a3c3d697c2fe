#!/bin/bash
mkdir -p gpurun_out
{
CUBEZ_STEP_LANE_MIN=4096 timeout 600 python tools/lanes_probe.py 32768 49152 2>&1
echo "--- 8192/12288 staggered: persistent (default) vs split + lanes"
timeout 600 python tools/lanes_probe.py 8192 12288 2>&1 | grep "lanes=1"
CUBEZ_STEP_LANE_MIN=2048 CUBEZ_FUSED_SPLIT=1 timeout 600 python tools/lanes_probe.py 12288 2>&1
} > gpurun_out/r02_lanes_mid.log 2>&1; cat gpurun_out/r02_lanes_mid.log
