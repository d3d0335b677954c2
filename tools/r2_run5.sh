#!/bin/bash
mkdir -p gpurun_out
T=./nitorch_fastmath_b200
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest5.log
timeout 300 $T/nfm_tune balance > gpurun_out/r2_balance5.log 2>&1
timeout 120 $T/nfm_tune_timeline timeline > gpurun_out/r2_timeline5.log 2>&1
echo "timeline rc=$?" >> gpurun_out/r2_timeline5.log
tail -3 gpurun_out/r2_pytest5.log
