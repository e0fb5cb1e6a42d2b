"""Prints the headline fields of a bench.py JSON line (gpurun log helper)."""
import json
import sys

L = [l for l in open(sys.argv[1]) if l.startswith("{")]
if not L:
    print("no JSON line in", sys.argv[1])
    sys.exit(0)
j = json.loads(L[-1])
r = j.get("roofline") or {}
print("impl", j.get("impl", "native"), "N", j.get("n_gpus"), "value", round(j["value"], 1), j["unit"], "step_ms", round(j["ms_per_step"], 4),
      "| scan_ms", r.get("avg_launch_ms"), "frac", r.get("frac"), "| e2e", round(j["e2e"]["value"], 1),
      "| e2e sync", ((j["e2e"].get("synchronous") or {}).get("value")),
      "| e2e new strings", (j["e2e"].get("new_strings_every_step") or {}).get("value"),
      "| cpu", (j.get("cpu_baseline") or {}).get("value"), "| launches/step", j.get("gpu_launches_per_step"),
      "| digest", j.get("result_digest"))
