#!/bin/bash
T=./nitorch_fastmath_b200
for v in r1 base static r1 base; do
  b=$T/nfm_ab_$v; [ $v = r1 ] && b=$T/nfm_tune_r1
  echo "== $v"; timeout 60 $b solve3 2>&1 | head -1; timeout 60 $b inv4d 2>&1 | head -3; timeout 60 $b det4d 2>&1 | head -2; timeout 60 $b solve4d 2>&1 | head -1
done > gpurun_out/r2_ab16.log
