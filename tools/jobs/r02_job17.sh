#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_job17_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_job17_tests.log
tail -4 gpurun_out/r02_job17_tests.log
timeout 600 python bench.py --no-k1 --no-configs > gpurun_out/r02_bench_b2shared.json 2> gpurun_out/r02_bench_b2shared.err; echo "bench rc $?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_bench_b2shared.json') if l.startswith('{')][-1]); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'rl', d['e2e_rl']['value']); print(json.dumps(d['strong'])[:400]); print(json.dumps(d.get('roofline',{}).get('ncu'))[:300])"
