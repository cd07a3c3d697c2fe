#!/bin/bash
# round-2 GPU job 4: whole GPU suite, the restructured bench (N = 1), ncu counters of this build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_job4_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_job4_tests.log
timeout 900 python bench.py > gpurun_out/r02_bench_v2.json 2> gpurun_out/r02_bench_v2.err; echo "bench rc $?" >> gpurun_out/r02_bench_v2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_v2.json 2>> gpurun_out/r02_bench_v2.err
bash tools/jobs/ncu_counters.sh
tail -4 gpurun_out/r02_job4_tests.log; tail -5 gpurun_out/r02_bench_v2.err; head -c 1500 gpurun_out/r02_bench_v2.json; echo; head -c 600 gpurun_out/r02_bench_ref_v2.json; tail -3 gpurun_out/ncu_fused.log
