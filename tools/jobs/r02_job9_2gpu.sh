#!/bin/bash
# round-2 GPU job 9 (2 GPUs): the in-library NCCL reduce (cz_run_*) and the strong / weak arms of bench.py at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_run_abi.py -m gpu -x -q > gpurun_out/r02_job9_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_job9_tests.log
tail -4 gpurun_out/r02_job9_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench rc $?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_bench_2gpu.json') if l.startswith('{')][-1]); print('value', d['value'], 'e2e', d['e2e']['value'], 'rl', d['e2e_rl']['value'], 'rl pipelined', d['e2e_rl_pipelined']['value']); print(json.dumps(d['strong']))"
tail -3 gpurun_out/r02_bench_2gpu.err
