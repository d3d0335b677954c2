#!/bin/bash
# round 2, GPU call 1: parity of the reworked tile kernel, launch timeline, equal-tile sweep, pool-kernel sweep
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
timeout 300 ./nitorch_fastmath_b200/nfm_tune_timeline timeline > gpurun_out/r2_timeline1.log 2>&1
timeout 300 ./nitorch_fastmath_b200/nfm_tune balance > gpurun_out/r2_balance1.log 2>&1
timeout 600 ./nitorch_fastmath_b200/nfm_tune pool > gpurun_out/r2_pool1.log 2>&1
tail -3 gpurun_out/r2_pytest1.log
