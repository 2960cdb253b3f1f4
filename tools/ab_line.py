"""Print the A/B-relevant numbers of one bench.py JSON line: label, ms/step, e2e, SM MHz, GEMM-family ms, attention ms, full-step ms."""
import json
import sys

label, path = sys.argv[1], sys.argv[2]
line = [x for x in open(path) if x.startswith("{")][-1]
d = json.loads(line)
r = d.get("roofline", {})
full = d.get("train_step_full") or {}
print(label, "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "MHz", d["clocks"].get("sm_mhz"), "gemm_ms %.2f" % r.get("gemm_ms_per_step", 0),
      "attn_ms %.2f" % r.get("attention", {}).get("ms_per_step", 0), "full_ms", full.get("ms_per_step"), "launches", d.get("gpu_launches_per_step"), flush=True)
