#!/bin/bash
# One GPU session under gpurun:  bash tools/gpu_session.sh [tests|bench|ncu ...]
#   tests : GPU tests -> smoke
#   bench : bench (N=1) -> reference arm
#   probe : kernel sweep (tools/gpu_probe.py, PROBE_FULL)
#   ncu   : launch list + full capture of the scan kernel (each only after the same command exited 0 without it)
#   ab    : A/B of kernel knobs
# Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
for what in "$@"; do
case "$what" in
tests)
  timeout 1500 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/gpu_tests.log | head -2
  grep -E "^FAILED|^ERROR|^E  " gpurun_out/gpu_tests.log | head -30
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
  ;;
tests_all)
  timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/gpu_tests.log | head -2
  grep -E "^FAILED|^ERROR|^E  " gpurun_out/gpu_tests.log | head -40
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
  ;;
bench)
  timeout 400 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"; tail -3 gpurun_out/bench_n1.err
  python tools/show_bench.py gpurun_out/bench_n1.json
  timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"
  python tools/show_bench.py gpurun_out/bench_ref.json
  ;;
probe)
  PROBE_FULL=1 timeout 400 python tools/gpu_probe.py 2>&1 | grep -E "timing|error|mismatch|match" > gpurun_out/probe.log; cut -c1-150 gpurun_out/probe.log
  ;;
ab)
  # A/B of the scan kernel's knobs on one box (tools/gpu_probe.py timings + event counters)
  for v in ${AB_VARIANTS:-"MPR_X=0" "MPR_NO_REGLIST=1" "MPR_DEBUG_COUNTERS=1"}; do
    echo "== $v"
    env ${v//,/ } PROBE_AB=1 timeout 300 python tools/gpu_probe.py 2>&1 | grep -E "timing|error|counters" | cut -c1-200
  done > gpurun_out/ab.log 2>&1; cat gpurun_out/ab.log
  ;;
sanitize_*)
  tool=${what#sanitize_}
  timeout 300 python tools/sanitize_target.py > gpurun_out/sanitize_plain.log 2>&1 &&
  timeout 1500 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_$tool.log python tools/sanitize_target.py > gpurun_out/sanitize_$tool.out 2>&1
  echo "sanitizer $tool exit=$?"; tail -3 gpurun_out/sanitize_plain.log; tail -5 gpurun_out/sanitizer_$tool.log
  ;;
ncu_cfg4)
  CMD="python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline --quick"
  timeout 400 $CMD > gpurun_out/plain_cfg4.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 4 -c 1 -o gpurun_out/prof_cfg4 -f $CMD > gpurun_out/ncu_cfg4.log 2>&1
  echo "ncu cfg4 exit=$?"; python tools/show_bench.py gpurun_out/plain_cfg4.log
  ;;
hostprof)
  timeout 300 python tools/profile_host.py 400 2>&1 | head -60
  ;;
timeline)
  for c in ${TIMELINE_CASES:-"128 1250000 512 5" "128 1250000 512 32"}; do
    MPR_DEBUG_COUNTERS=2 timeout 120 python tools/probe_one.py $c
  done > gpurun_out/timeline.log 2>&1; cat gpurun_out/timeline.log
  ;;
ncu_k32)
  CMD="python tools/probe_one.py 128 1250000 512 32"
  timeout 300 $CMD > gpurun_out/plain_k32.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 2 -c 1 -o gpurun_out/prof_k32 -f $CMD > gpurun_out/ncu_k32.log 2>&1
  echo "ncu k32 exit=$?"; tail -2 gpurun_out/plain_k32.log
  ;;
ncu)
  CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --quick"
  timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
  echo "ncu launches exit=$?"
  timeout 300 $CMD > gpurun_out/plain2.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 2 -o gpurun_out/prof_scan -f $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit=$?"
  ;;
*) echo "unknown step $what";;
esac
done
