"""One-screen summary of a bench.py JSON line."""
import json
import sys

d = json.load(open(sys.argv[1]))
print("value %.1f (%.3f ms/step, host cpu %.3f)  e2e %.1f (%.3f ms/step, host cpu %.3f)  per-view %s  launches %s" % (
    d["value"], d["ms_per_step"], d.get("host_cpu_ms_per_step") or -1, d["e2e"]["value"], d["e2e"]["ms_per_step"],
    d["e2e"].get("host_cpu_ms_per_step") or -1, d.get("per_view_api", {}).get("value"), d.get("gpu_launches")))
if d.get("stages"):
    print({k: round(v["ms_avg"], 4) for k, v in d["stages"].items() if v["ms_avg"]})
for k in ("allreduce_check", "exchange_check"):
    if d.get(k):
        print(k, d[k])
for c, v in (d.get("configs") or {}).items():
    print(c, {k: (round(x, 1) if isinstance(x, float) else x) for k, x in v.items() if k != "workload"})
