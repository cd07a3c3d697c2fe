#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4880 -c 800 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 100 --warmup 5 --no-k1 --no-configs --no-strong --e2e-steps 2 > gpurun_out/r02_launches_bench.log 2>&1
grep -c k_world_fused gpurun_out/r02_launches_bench.csv; tail -2 gpurun_out/r02_launches_bench.log | cut -c1-300
