#!/bin/bash
# tests + N-GPU bench (args: N)
set -u
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/gpu_tests.log
  timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
  timeout 300 python bench.py --no-graph --no-cpu-baseline > gpurun_out/bench_n1_nograph.json 2>/dev/null; cat gpurun_out/bench_n1_nograph.json | cut -c1-400
else
  for n in $(echo $N | tr ',' ' '); do
    timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "N=$n exit=$?"
    cat gpurun_out/bench_n$n.json | cut -c1-2600; tail -3 gpurun_out/bench_n$n.err | cut -c1-300
  done
fi
