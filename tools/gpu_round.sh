#!/bin/bash
# One GPU session: tests, smoke, bench (N=1), ncu launch list + full capture of the scan kernel.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit=$?" | tee -a gpurun_out/gpu_tests.log
tail -3 gpurun_out/gpu_tests.log
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"
cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"
cat gpurun_out/bench_ref.json
if [ "${1:-}" = "ncu" ]; then
  CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
  timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
  echo "ncu launches exit=$?"
  timeout 300 $CMD > gpurun_out/plain2.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 2 -o gpurun_out/prof_scan -f $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit=$?"
  tail -3 gpurun_out/ncu_full.log
fi
