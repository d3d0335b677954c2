#!/bin/bash
mkdir -p gpurun_out
T=./nitorch_fastmath_b200
timeout 200 $T/nfm_tune balance > gpurun_out/r2_balance4.log 2>&1
timeout 120 $T/nfm_tune_timeline timeline > gpurun_out/r2_timeline4.log 2>&1
echo "timeline rc=$?" >> gpurun_out/r2_timeline4.log
for w in solve3 solve6 invert6 solve10 inv4d solve4d; do
  timeout 120 $T/nfm_tune $w >> gpurun_out/r2_big4.log 2>&1
done
