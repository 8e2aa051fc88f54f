#!/usr/bin/env python
"""Prints the headline numbers and the per-launch table of bench.py JSON lines (developer helper)."""
import json
import sys
for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print(f, "value", round(d["value"]), "img/s  ms/step", round(d["ms_per_step"], 3), " e2e", e.get("value") and round(e["value"]), d.get("clocks"))
    for s in d.get("stages") or []:
        print("  %-14s %8.1f us  %7.1f MB  frac %.3f" % (s["stage"], s["us"], s["alg_MB"], s["frac"]))
