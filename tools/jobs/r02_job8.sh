#!/bin/bash
# round-2 GPU job 8: adjacency lists in global memory (26 k-contact capacity), pipelined RL step, forces; the 600-frame cfg3 run
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_forces.py tests/test_gpu_rl_step.py tests/test_gpu_islands.py tests/test_gpu_broadphase.py tests/test_gpu_cfg3.py tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/r02_job8_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_job8_tests.log
tail -4 gpurun_out/r02_job8_tests.log
timeout 900 python tools/cfg3_probe.py --frames 600 --block 50 > gpurun_out/r02_cfg3_600.log 2>&1
tail -14 gpurun_out/r02_cfg3_600.log
timeout 300 python bench.py --no-configs --no-k1 --no-strong --steps 200 > gpurun_out/r02_bench_rlasync.json 2> gpurun_out/r02_bench_rlasync.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_rlasync.json').read().strip().splitlines()[-1]); print('value', d['value'], 'e2e', d['e2e']['value'], 'rl', d['e2e_rl']['value'], 'rl pipelined f32', d['e2e_rl_pipelined']['value'])"
