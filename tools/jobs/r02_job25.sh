#!/bin/bash
mkdir -p gpurun_out
{
for cs in 3 4 6; do echo "--- comp streams $cs"; CUBEZ_HOST_COMP_STREAMS=$cs python tools/rl_probe.py 3 4 6 8 12 2>&1 | tail -5; done
} > gpurun_out/r02_rl_sync_streams.log 2>&1; cat gpurun_out/r02_rl_sync_streams.log
