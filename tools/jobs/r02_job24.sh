#!/bin/bash
mkdir -p gpurun_out
{
echo "--- staggered, 8192 / 10240 worlds: persistent vs split in 4 lanes"
for W in 8192 10240; do
python tools/lanes_probe.py $W 2>&1 | grep "lanes=1"
CUBEZ_FUSED_SPLIT=1 python tools/lanes_probe.py $W 2>&1 | grep "lanes=4"
done
echo "--- from t = 0, 600 frames, 8192 worlds: persistent vs split in 4 lanes"
python tools/strong_probe.py 8192 2>&1 | head -2
} > gpurun_out/r02_lanes_8192.log 2>&1; cat gpurun_out/r02_lanes_8192.log
