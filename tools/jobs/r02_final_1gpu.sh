#!/bin/bash
# round-2 final 1-GPU job: whole GPU suite, smoke, bench (ours + reference arm), ncu counters and launch list of the committed code
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_final_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1
bash tools/jobs/ncu_counters.sh
python tools/ncu_counters.py --fused gpurun_out/ncu_fused.csv --k1 gpurun_out/ncu_k1.csv --k2 gpurun_out/ncu_k2_all.csv > gpurun_out/ncu_counters.log 2>&1
cp profiles/ncu_counters.json gpurun_out/ncu_counters.json
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc $?" >> gpurun_out/r02_bench_final.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref_final.json 2>> gpurun_out/r02_bench_final.err
# launch list of the TIMED region: the 600-frame pre-roll is 4 800 launches (two lanes of order + A + B + C per frame) after ~30 set-up
# kernels; skipped launches run unprofiled, the 800 captured ones are the 100 timed frames (shares agree with the bench line, absolutes
# are ncu's serialised cold-cache times)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4880 -c 800 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 100 --warmup 5 --no-k1 --no-configs --no-strong --e2e-steps 2 > gpurun_out/r02_launches_bench.log 2>&1
tail -4 gpurun_out/r02_final_tests.log; cat gpurun_out/r02_final_smoke.log | tail -2; tail -3 gpurun_out/r02_bench_final.err; head -c 700 gpurun_out/r02_bench_final.json; echo; cat gpurun_out/ncu_counters.json
