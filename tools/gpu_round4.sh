#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/gpu_tests.log | head -2
timeout 200 python tools/gpu_probe.py 2>&1 | grep -E "timing|error" > gpurun_out/probe4.log; cat gpurun_out/probe4.log | cut -c1-150
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --bank-rows 1250000"
timeout 200 $CMD > gpurun_out/plain_s8.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/prof_scan_shard8 -f $CMD > gpurun_out/ncu_s8.log 2>&1
echo "ncu exit=$?"
