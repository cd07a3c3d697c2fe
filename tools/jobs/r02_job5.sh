#!/bin/bash
# round-2 GPU job 5: velocity-response precompute in the fused loops: parity (whole suite) and A/B timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_job5_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_job5_tests.log
for pre in 1 0; do
  CUBEZ_FUSED_VEL_PRE=$pre timeout 600 python bench.py --no-configs --no-k1 --no-strong --steps 200 > gpurun_out/r02_bench_velpre$pre.json 2> gpurun_out/r02_bench_velpre$pre.err
done
tail -4 gpurun_out/r02_job5_tests.log
for pre in 1 0; do python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_velpre$pre.json').read().strip().splitlines()[-1]); print('velpre $pre', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_rl']['value'])"; done
