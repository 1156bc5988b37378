"""Prints the interesting fields of a bench.py JSON line: python tools/print_bench.py gpurun_out/bench.json"""
import json
import sys

l = json.loads(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bench.json").read().strip().splitlines()[-1])
print("exact", round(l["value"]), "perm/s", round(l["ms_per_step"], 3), "ms  e2e", round(l["e2e"]["value"]), "frac", round(l["roofline"]["frac"], 3), l["roofline"].get("kernel_ms"))
if l.get("fast"):
    print("fast ", round(l["fast"]["value"]), "perm/s", round(l["fast"]["ms_per_step"], 3), "ms  e2e", round(l["fast"]["e2e"]["value"]), "frac", round(l["roofline_fast"]["frac"], 3))
if l.get("nhood_cfg5_strong"):
    print("cfg5 ", round(l["nhood_cfg5_strong"]["value"]), "perm/s")
m = l.get("moran") or {}
if m.get("value"):
    print("moran", round(m["value"]), "genes/s", round(m["ms_per_step"], 3), "ms  e2e s", m.get("e2e", {}).get("seconds_all"), "n_perms_100", m.get("n_perms_100_seconds"), "frac", round(m.get("roofline", {}).get("frac", 0), 4),
          "parity", m.get("parity_max_abs_err_sample"))
for k in ("co_occurrence", "ripley_L"):
    d = l.get(k) or {}
    if d.get("kernel_ms"):
        print(k, "kernel ms", round(d["kernel_ms"], 2), "e2e s", d.get("e2e", {}).get("seconds_all"))
print("cpu", l.get("cpu_baseline"), "clocks", l.get("clocks"))
