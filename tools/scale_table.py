#!/usr/bin/env python
"""Condense bench.py JSON lines (one per run) into a scaling table."""
import json
import sys

rows = [json.loads(l) for p in sys.argv[1:] for l in open(p) if l.startswith("{")]
base = {}
print("%-46s %5s %6s %12s %10s %9s %7s %10s %12s" % ("workload", "gpus", "steps", "Gmat/s", "us/step", "GB/s/GPU", "frac", "efficiency", "e2e Gmat/s"))
for d in rows:
    c, r = d["config"], d["roofline"]
    key = (c["workload"], d["steps"])
    if d["n_gpus"] == 1:
        base[c["workload"]] = d["value"]
    eff = d["value"] / (base[c["workload"]] * d["n_gpus"]) if c["workload"] in base else float("nan")
    e2e = d.get("e2e") or {}
    print("%-46s %5d %6d %12.2f %10.2f %9.0f %7.3f %10.3f %12s" % (
        c["workload"][:46], d["n_gpus"], d["steps"], d["value"] / 1e9, d["ms_per_step"] * 1e3, r["achieved"], r["frac"], eff,
        "%.2f" % (e2e["value"] / 1e9) if e2e else "-"))
