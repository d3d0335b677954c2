#!/bin/bash
# fp64 dense inverse, orders 8..10: two lanes per matrix (default) vs one thread per matrix (NFM_DISABLE_PAIR_INVERSE=1)
for n in 8 9 10; do
  for off in 0 1; do
    NFM_DISABLE_PAIR_INVERSE=$off python bench.py --kind batch_inv --n $n --dtype f64 --batch 4194304 --steps 20 --warmup 5 --no-e2e --no-cpu 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batchinv f64 n=$n pair_disabled=$off: %8.1f us  %6.0f GB/s  frac %.3f  %s' % (d['ms_per_step']*1e3, d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['kernel'][:40]))"
  done
done
