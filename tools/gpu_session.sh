#!/bin/bash
# One GPU session under gpurun:  bash tools/gpu_session.sh [ncu]
#   GPU tests -> smoke -> bench (N=1) -> reference arm -> sweep; with "ncu": launch list + full capture of the scan kernel
#   (each ncu run only after the same command exited 0 without it).  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/gpu_tests.log | head -2
grep -E "^FAILED|^E  " gpurun_out/gpu_tests.log | head -12
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"
python - <<'PY'
import json
L = [l for l in open("gpurun_out/bench_n1.json") if l.startswith("{")]
if L:
    j = json.loads(L[-1]); r = j["roofline"]
    print("value", round(j["value"]), "q/s  step_ms", round(j["ms_per_step"], 4), " scan_ms", round(r["avg_launch_ms"], 4),
          "frac", round(r["frac"], 3), " e2e", round(j["e2e"]["value"]), "q/s  cpu", round(j.get("cpu_baseline", {}).get("value", 0), 1))
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"
PROBE_FULL=1 timeout 200 python tools/gpu_probe.py 2>&1 | grep -E "timing|error" > gpurun_out/probe.log; cut -c1-130 gpurun_out/probe.log
if [ "${1:-}" = "ncu" ]; then
  CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
  timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
  echo "ncu launches exit=$?"
  timeout 300 $CMD > gpurun_out/plain2.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 2 -o gpurun_out/prof_scan -f $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit=$?"
fi
