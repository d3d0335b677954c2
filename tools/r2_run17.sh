#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/sweep.sh > gpurun_out/r2_final_workloads.log 2>&1; cat gpurun_out/r2_final_workloads.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
