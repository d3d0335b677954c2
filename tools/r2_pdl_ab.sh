#!/bin/bash
# what PDL + the L2 prefetch ahead of griddepcontrol.wait buy on a 2M-matrix slab (config 2 at 8 GPUs), bench.py, 200 steps
mkdir -p gpurun_out
for env in "" "NFM_DISABLE_PDL=1"; do
  for i in 1 2; do
    env $env python bench.py --batch 2097152 --steps 200 --warmup 20 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('${env:-PDL + prefetch (default)}: %.2f us/step  %.0f GB/s' % (d['ms_per_step'] * 1e3, d['roofline']['achieved']))"
  done
done
python tools/launch_overhead.py
