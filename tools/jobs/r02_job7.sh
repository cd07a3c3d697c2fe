#!/bin/bash
# round-2 GPU job 7: force/torque input, islands (fixed test), register-budget A/B of the loop phases, 600-frame cfg3 run
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_forces.py tests/test_gpu_islands.py tests/test_gpu_rl_step.py tests/test_gpu_worlds.py -m gpu -x -q > gpurun_out/r02_job7_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_job7_tests.log
for mb in 3 4; do
  CUBEZ_FUSED_SPLIT_MINB=$mb timeout 600 python bench.py --no-configs --no-k1 --no-strong --steps 200 > gpurun_out/r02_bench_minb$mb.json 2> gpurun_out/r02_bench_minb$mb.err
done
tail -4 gpurun_out/r02_job7_tests.log
for mb in 3 4; do python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_minb$mb.json').read().strip().splitlines()[-1]); print('minb $mb', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_rl']['value'])"; done
timeout 1500 python tools/cfg3_probe.py --frames 600 --block 25 > gpurun_out/r02_cfg3_600.log 2>&1
tail -8 gpurun_out/r02_cfg3_600.log
