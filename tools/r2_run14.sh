#!/bin/bash
mkdir -p gpurun_out
./nitorch_fastmath_b200/nfm_ab_base reps20 > gpurun_out/r2_reps20.log 2>&1; cat gpurun_out/r2_reps20.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench14.json 2>gpurun_out/r2_bench14.err; tail -c 300 gpurun_out/r2_bench14.err
for k in "sym_invert --n 3" "sym_invert --n 1" "batch_inv --n 1" "batch_det --n 1" "sym_solve --n 1" "batch_inv --n 2"; do
 python bench.py --kind $k --dtype f32 --batch 8388608 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('%s: frac %.3f' % (d['config']['workload'], d['roofline']['frac']))"
done
