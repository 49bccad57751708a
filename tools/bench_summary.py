#!/usr/bin/env python
"""Print a one-line summary for every bench.py JSON line on stdin (label = argv[1])."""
import json
import sys

label = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline", {})
    print(label, "ms/step %.4f value %.4g warmL2 %.4g plan %.4g e2e %.4g (%.3f ms) frac %.3f clocks %s cpu %s" % (
        d["ms_per_step"], d["value"], d.get("value_warm_l2_rank0", 0),
        d.get("plan_persistent_rank0", {}).get("value", 0), d["e2e"]["value"],
        d["e2e"].get("ms_per_step", 0), r.get("frac", 0), d.get("clocks", {}).get("sm_mhz"),
        d.get("cpu_baseline", {}).get("value")))
