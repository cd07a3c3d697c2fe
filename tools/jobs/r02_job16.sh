#!/bin/bash
mkdir -p gpurun_out
{
run() { echo "--- weights=$1 splitmin=$2 comp=$3"; CUBEZ_HOST_CHUNK_WEIGHTS=$1 CUBEZ_HOST_SPLIT_MIN=$2 CUBEZ_HOST_COMP_STREAMS=$3 python tools/e2e_probe.py 0 2>&1 | tail -1; }
echo "--- default 12"; python tools/e2e_probe.py 12 2>&1 | tail -1
run 0.1,0.2,0.4,0.8,1,1,1,1,1,1,0.7,0.4 0 3
run 0.1,0.2,0.4,0.8,1,1,1,1,1,1,0.7,0.4 4000 3
run 0.1,0.2,0.4,0.8,1,1,1,1,1,1,0.7,0.4 100000 3
run 0.1,0.2,0.4,0.8,1,1,1,1,1,1,0.7,0.4 4000 4
run 0.15,0.3,0.6,1,1,1,1,1,1,1,0.6,0.3 6000 3
run 0.15,0.3,0.6,1,1,1,1,1,1,1,0.6,0.3 6000 6
run 0.25,0.5,1,1,1,1,1,1,0.5,0.25 6000 3
run 0.1,0.15,0.25,0.4,0.6,0.8,1,1,1,1,1,1,1,0.8,0.5,0.3 6000 4
run 0.3,0.6,1,1,1,1,1,0.6 6000 3
echo "--- trace"; CUBEZ_HOST_TRACE=1 CUBEZ_HOST_CHUNK_WEIGHTS=0.1,0.2,0.4,0.8,1,1,1,1,1,1,0.7,0.4 CUBEZ_HOST_SPLIT_MIN=4000 python tools/e2e_probe.py 0 2>&1 | tail -15
} > gpurun_out/r02_e2e_ramp.log 2>&1
cat gpurun_out/r02_e2e_ramp.log
