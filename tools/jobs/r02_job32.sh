#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/fuzz_isolate.py > gpurun_out/r02_fuzz_isolate.log 2>&1; cat gpurun_out/r02_fuzz_isolate.log
