#!/bin/bash
# tools/sweep.sh [extra bench.py args] -- one bench.py line per workload, condensed.
for w in sym_solve3 sym_matvec3 sym_solve6 sym_invert6 sym_solve10 sym_solve3_1m dense_inv4_f64 dense_det4_f64 dense_solve4_f64; do
  timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu --no-e2e "$@" 2>&1 | W=$w python -c '
import sys, json, os
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception:
        print(l.rstrip()); continue
    r = d["roofline"]
    print("%-18s %8.2f Gmat/s %7.1f GB/s  frac %.3f  %8.1f us  launches %d  clk %s" % (
        os.environ["W"], d["value"] / 1e9, r["achieved"], r["frac"], r["avg_launch_us"], d["gpu_launches"], d["clocks"]["sm_mhz"]))
'
done
