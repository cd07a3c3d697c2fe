#!/bin/bash
mkdir -p gpurun_out
python tools/pcie_pattern_probe.py > gpurun_out/r02_pcie_pattern.log 2>&1; cat gpurun_out/r02_pcie_pattern.log
