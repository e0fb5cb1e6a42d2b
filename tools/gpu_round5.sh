#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/gpu_tests.log | head -2; grep -E "FAILED|Error" gpurun_out/gpu_tests.log | head
timeout 200 python tools/gpu_probe.py 2>&1 | grep -E "timing|error" > gpurun_out/probe5.log; cut -c1-150 gpurun_out/probe5.log
MPR_NO_CLUSTER=1 timeout 200 python tools/gpu_probe.py 2>&1 | grep -E "timing|error" > gpurun_out/probe5_nocluster.log; cut -c1-110 gpurun_out/probe5_nocluster.log
