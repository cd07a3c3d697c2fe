#!/bin/bash
# ncu --set full of the committed code: the three fused phase launches of one frame (one lane), the large-world resolver on the settled cfg3 frame
mkdir -p gpurun_out
CUBEZ_STEP_LANES=1 PROF_WORLDS=65536 PROF_FRAMES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_world_fused --launch-skip 1803 --launch-count 3 -f -o gpurun_out/r02_fused_phases_v7 python tools/profile_fused.py > gpurun_out/r02_fused_phases_v7_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resolve -c 1 -f -o gpurun_out/r02_resolve_settled_v5 python tools/cfg3_probe.py --load tests/golden/pile4096_f100.npz --steps 1 > gpurun_out/r02_resolve_ncu_v5.log 2>&1
ls -la gpurun_out/*.ncu-rep
