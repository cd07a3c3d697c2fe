#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/probes/_memcpy_batch_probe > gpurun_out/r02_memcpy_batch_probe.log 2>&1; cat gpurun_out/r02_memcpy_batch_probe.log
