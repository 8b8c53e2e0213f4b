#!/usr/bin/env python
"""Print the interesting numbers of bench.py JSON lines read from stdin (one summary line each)."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline", {})
    print(" ".join(str(x) for x in (
        sys.argv[1] if len(sys.argv) > 1 else "", d["config"].get("workload", "")[:40],
        "ms=%.2f" % d["ms_per_step"], "e2e_ms=%.2f" % d.get("e2e", {}).get("ms_per_step", -1),
        "kern_ms=%.2f" % r.get("kernel_ms", -1), "grid_ms=%.2f" % r.get("gridlink_ms", -1),
        "frac=%.4f" % r.get("frac", -1), "lev=%.2f" % r.get("levels_per_eval", -1),
        "neval=%.3e" % r.get("n_eval", -1), "nana=%.3e" % r.get("n_analytic", -1), "jobs=%.3e" % r.get("n_jobs", -1),
        "tiles=%d" % r.get("n_tiles", -1), "lat=%s" % d["config"].get("device_lattice"))))
