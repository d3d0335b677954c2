#!/bin/bash
mkdir -p gpurun_out
T=./nitorch_fastmath_b200
timeout 120 $T/nfm_tune_timeline timeline > gpurun_out/r2_timeline2.log 2>&1
timeout 120 $T/nfm_tune_timeline_rd timeline > gpurun_out/r2_timeline2_rd.log 2>&1
: > gpurun_out/r2_pool2.log
for b in pool8d poolx8d_w9 poolx8d_w10 poolx8d_w12 pool10d pool10f pool8f pool6d pool4d pooldet10d poolsolve10d poolsolve10f poolsymlu10f; do
  timeout 200 $T/nfm_tune $b >> gpurun_out/r2_pool2.log 2>&1
done
which compute-sanitizer > gpurun_out/r2_sanitizer.log 2>&1
timeout 300 compute-sanitizer --tool memcheck $T/nfm_tune poolx8d_w10 >> gpurun_out/r2_sanitizer.log 2>&1
tail -5 gpurun_out/r2_sanitizer.log
