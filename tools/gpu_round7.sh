#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/gpu_tests.log | head -2; grep -E "^FAILED|^E  " gpurun_out/gpu_tests.log | head -12
timeout 150 python tools/gpu_probe.py 2>&1 | grep -E "timing|error" > gpurun_out/probe7_s2.log; cut -c1-118 gpurun_out/probe7_s2.log
MPR_STAGE_SUBS=1 timeout 150 python tools/gpu_probe.py 2>&1 | grep -E "timing|error" > gpurun_out/probe7_s1.log; cut -c1-118 gpurun_out/probe7_s1.log
