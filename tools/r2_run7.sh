#!/bin/bash
mkdir -p gpurun_out
T=./nitorch_fastmath_b200
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log
bash tools/sweep.sh > gpurun_out/r2_sweep7.log 2>&1
: > gpurun_out/r2_pool7.log
for b in pool_inv5d pool_inv4d pool_inv7f pool_inv6f pool_inv5f pool_solve6d pool_inv8d pool_inv6d pool_inv8f pool_symlu10f; do
  timeout 300 $T/nfm_tune $b 2>&1 | grep -B1 "rule" >> gpurun_out/r2_pool7.log
done
tail -3 gpurun_out/r2_pytest7.log; cat gpurun_out/r2_sweep7.log
