#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/strong_probe.py > gpurun_out/r02_strong_probe.log 2>&1; cat gpurun_out/r02_strong_probe.log
