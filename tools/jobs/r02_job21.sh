#!/bin/bash
mkdir -p gpurun_out
{
for min in 8192 4096; do echo "--- lane min $min"; CUBEZ_STEP_LANE_MIN=$min CUBEZ_FUSED_SPLIT=1 timeout 600 python tools/lanes_probe.py 16384 24576 8192 2>&1; done
echo "--- strong-style (t=0, 600 frames) W=16384: lanes 1 vs 2"
CUBEZ_STEP_LANES=1 python tools/strong_probe.py 16384 2>&1 | head -1
CUBEZ_STEP_LANE_MIN=8192 python tools/strong_probe.py 16384 2>&1 | head -1
echo "--- W=8192 split + 2 lanes of 4096 vs persistent"
CUBEZ_STEP_LANE_MIN=4096 CUBEZ_FUSED_SPLIT=1 python tools/strong_probe.py 8192 2>&1 | head -1
python tools/strong_probe.py 8192 2>&1 | head -1
} > gpurun_out/r02_lanes_small.log 2>&1; cat gpurun_out/r02_lanes_small.log
