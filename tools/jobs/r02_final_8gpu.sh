#!/bin/bash
# round-2 final multi-GPU job: in-library NCCL reduce over all devices, bench at N = $1 (weak + strong arms)
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_run_abi.py -m gpu -q > gpurun_out/r02_run_abi_${N}gpu.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_run_abi_${N}gpu.log
tail -3 gpurun_out/r02_run_abi_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 5 --no-k1 --no-configs > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench rc $?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_bench_${N}gpu.json') if l.startswith('{')][-1]); print('value', d['value'], 'e2e', d['e2e']['value'], 'rl', d['e2e_rl']['value'], 'rl pipelined', d['e2e_rl_pipelined']['value']); print(json.dumps(d['strong']))"
