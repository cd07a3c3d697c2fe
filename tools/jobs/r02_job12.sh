#!/bin/bash
mkdir -p gpurun_out
{
python tools/pcie_probe.py
echo "--- default (6 chunks, split)"; CUBEZ_HOST_TRACE=1 python tools/e2e_probe.py 6 2>&1 | tail -12
echo "--- no split"; CUBEZ_HOST_NO_SPLIT=1 CUBEZ_HOST_TRACE=1 python tools/e2e_probe.py 6 8 2>&1 | tail -24
echo "--- 8/10/12 chunks split"; python tools/e2e_probe.py 8 10 12 2>&1 | tail -3
} > gpurun_out/r02_e2e_trace.log 2>&1
tail -60 gpurun_out/r02_e2e_trace.log
