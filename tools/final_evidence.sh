#!/bin/bash
# One-GPU evidence pass (run under gpurun): tests, driver-style bench + reference arm, other workloads, A/B, timelines.
set -u
mkdir -p gpurun_out
bash tools/gpu_session.sh tests_all
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"
python tools/show_bench.py gpurun_out/bench_n1.json
timeout 400 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"
python tools/show_bench.py gpurun_out/bench_ref.json
for w in cfg3 cfg4; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w exit=$?"
  python tools/show_bench.py gpurun_out/bench_$w.json
done
timeout 300 python bench.py --bank-rows 1250000 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_shard8.json 2> gpurun_out/bench_shard8.err
python tools/show_bench.py gpurun_out/bench_shard8.json
# A/B of the scan's design choices, each in a fresh process per shape (a long probe session runs power-capped)
for v in MPR_X=0 MPR_STATIC_TILES=1 MPR_NO_GTHR=1 MPR_FIRST_WAIT_NS=-1 MPR_NO_QCOOP=1 MPR_NO_TAIL_FLOOR=1 MPR_NO_REGLIST=1; do
  for c in "128 1250000 512 5" "128 1250000 512 16" "128 1250000 512 32"; do
    echo -n "$v | "; env $v timeout 120 python tools/probe_one.py $c 2>&1 | head -1 | cut -c1-50
  done
done > gpurun_out/ab.log 2>&1
cat gpurun_out/ab.log
for c in "128 1250000 512 5" "128 1250000 512 32" "16 1062912 1024 5" "128 1000000 1024 5"; do
  MPR_DEBUG_COUNTERS=2 timeout 120 python tools/probe_one.py $c
done > gpurun_out/timeline.log 2>&1
for c in "17 1000000 1024 5" "64 1000000 1024 5" "64 1000000 1024 16" "96 1000000 1024 5" "128 1000000 1024 5" "256 1000000 1024 5" "1024 1000000 1024 5" "16 1000000 1024 16" "1024 1048576 512 5"; do
  timeout 120 python tools/probe_one.py $c 2>&1 | head -1 | cut -c1-50
done > gpurun_out/shapes.log 2>&1
cat gpurun_out/shapes.log
