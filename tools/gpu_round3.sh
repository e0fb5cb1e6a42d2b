#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/gpu_tests.log
timeout 200 python tools/gpu_probe.py 2>&1 | grep -E "timing|error" > gpurun_out/probe3.log; cat gpurun_out/probe3.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; echo "bench exit=$?"; cut -c1-200 gpurun_out/bench_n1b.json; python -c "
import json; j=json.loads([l for l in open('gpurun_out/bench_n1b.json') if l.startswith('{')][-1]); print('e2e', j['e2e']['value'], j['e2e']['ms_per_step'], 'roof', j['roofline']['frac'], j['roofline']['launch_ms_min_median_max'])"
timeout 200 python bench.py --no-cpu-baseline --bank-rows 1250000 > gpurun_out/bench_shard8.json 2>/dev/null; python -c "
import json; j=json.loads([l for l in open('gpurun_out/bench_shard8.json') if l.startswith('{')][-1]); print('1.25M rows: value', j['value'], 'step_ms', j['ms_per_step'], 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], 'roof', j['roofline']['frac'], j['roofline']['launch_ms_min_median_max'])"
