#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest10.log
bash tools/sweep.sh > gpurun_out/r2_sweep10.log 2>&1
tail -5 gpurun_out/r2_pytest10.log; cat gpurun_out/r2_sweep10.log
