#!/bin/bash
# tools/sweep_all_n.sh -- roofline fraction for every order 1..10, both scalar
# types, for the main routines (device-resident, 4M..16M matrices each)
for dt in ${DTYPES:-f32 f64}; do for kind in ${KINDS:-sym_solve sym_matvec sym_invert batch_inv batch_det batch_solve}; do for n in ${ORDERS:-1 2 3 4 5 6 7 8 9 10}; do
  timeout 120 python bench.py ${EXTRA_ARGS:-} --kind $kind --n $n --dtype $dt --batch 8388608 --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | K="$kind n=$n $dt" python -c '
import sys, json, os
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception:
        print(os.environ["K"], "ERR", l.rstrip()[:150]); continue
    r = d["roofline"]
    print("%-24s %8.2f Gmat/s %7.1f GB/s  frac %.3f  %8.1f us" % (os.environ["K"], d["value"] / 1e9, r["achieved"], r["frac"], r["avg_launch_us"]))
'
done; done; done
