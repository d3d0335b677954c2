#!/bin/bash
mkdir -p gpurun_out
T=./nitorch_fastmath_b200
for i in 1 2; do
for w in solve3 inv4d solve4d; do
  echo "== r1 $w" ; timeout 120 $T/nfm_tune_r1 $w 2>&1
  echo "== r2 $w" ; timeout 120 $T/nfm_tune $w 2>&1
done
done > gpurun_out/r2_ab8.log
