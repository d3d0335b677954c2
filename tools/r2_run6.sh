#!/bin/bash
mkdir -p gpurun_out
T=./nitorch_fastmath_b200
for b in geo_solve3 geo_matvec3 geo_solve6 geo_invert6 geo_solve10 geo_solve4 geo_solve3d; do
  timeout 300 $T/nfm_tune $b > gpurun_out/r2_$b.log 2>&1
done
: > gpurun_out/r2_pool6.log
for b in pool_inv8d pool_inv10d pool_inv6d pool_inv10f pool_inv8f pool_det10d pool_solve10d pool_solve8d pool_solve10f pool_symlu10f pool_symlu10d; do
  timeout 300 $T/nfm_tune $b >> gpurun_out/r2_pool6.log 2>&1
done
wc -l gpurun_out/r2_geo_*.log gpurun_out/r2_pool6.log
