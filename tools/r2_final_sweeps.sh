#!/bin/bash
# final sweeps of round 2: every order x routine x dtype, the workload list, small-step protocol check
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest13.log
tail -3 gpurun_out/r2_pytest13.log
for s in 20 200; do for i in 1 2 3; do
  python bench.py --batch 2097152 --steps $s --warmup 5 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('2M-matrix slab, steps %d: %.2f us/step  frac %.3f clocks %s' % (d['steps'], d['ms_per_step'] * 1e3, d['roofline']['frac'], d['clocks']))"
done; done > gpurun_out/r2_small_step_protocol.log 2>&1
cat gpurun_out/r2_small_step_protocol.log
bash tools/sweep.sh > gpurun_out/r2_final_workloads.log 2>&1; cat gpurun_out/r2_final_workloads.log
DTYPES=f32 bash tools/sweep_all_n.sh > gpurun_out/r2_sweep_all_orders_f32.log 2>&1
DTYPES=f64 bash tools/sweep_all_n.sh > gpurun_out/r2_sweep_all_orders_f64.log 2>&1
wc -l gpurun_out/r2_sweep_all_orders_f*.log
