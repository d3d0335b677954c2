#!/bin/bash
mkdir -p gpurun_out
T=./nitorch_fastmath_b200
timeout 120 $T/nfm_tune balance > gpurun_out/r2_balance3_pf.log 2>&1
timeout 120 $T/nfm_tune_nopf balance > gpurun_out/r2_balance3_nopf.log 2>&1
timeout 120 $T/nfm_tune_timeline timeline > gpurun_out/r2_timeline3.log 2>&1
for w in solve3 solve6 invert6 solve10 inv4d solve4d; do
  timeout 120 $T/nfm_tune $w >> gpurun_out/r2_big3_pf.log 2>&1
  timeout 120 $T/nfm_tune_nopf $w >> gpurun_out/r2_big3_nopf.log 2>&1
done
: > gpurun_out/r2_pool3.log
for b in pool_inv8d pool_inv10d pool_inv6d pool_inv10f pool_inv8f pool_det10d pool_solve10d pool_solve8d pool_solve10f pool_symlu10f pool_symlu10d; do
  timeout 300 $T/nfm_tune $b >> gpurun_out/r2_pool3.log 2>&1
done
grep -c . gpurun_out/r2_pool3.log
