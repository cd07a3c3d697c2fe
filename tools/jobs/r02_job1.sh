#!/bin/bash
# round-2 GPU job 1: whole GPU suite (after the ADVICE fixes), cfg3 at full size (timing + snapshot + ncu of the resolver), bench baseline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_job1_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_job1_tests.log
timeout 400 python tools/cfg3_probe.py --frames 110 --save gpurun_out/pile4096_f100.npz --at 100 > gpurun_out/r02_cfg3_probe.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resolve -c 1 -f -o gpurun_out/r02_resolve_settled python tools/cfg3_probe.py --load gpurun_out/pile4096_f100.npz --steps 1 > gpurun_out/r02_resolve_ncu.log 2>&1
timeout 600 python bench.py --steps 100 > gpurun_out/r02_bench_baseline.json 2> gpurun_out/r02_bench_baseline.err
tail -5 gpurun_out/r02_job1_tests.log; tail -8 gpurun_out/r02_cfg3_probe.log; tail -3 gpurun_out/r02_resolve_ncu.log; head -c 600 gpurun_out/r02_bench_baseline.json
