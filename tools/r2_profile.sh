#!/bin/bash
# ncu captures of round 2 (one gpurun call): launch list of the default bench command, and
# --set full captures of the dominant kernel of each headline workload and of the pool kernel.
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e"
$B > gpurun_out/r2_plain_default.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_default.csv $B > gpurun_out/r2_ncu_launches.log 2>&1
cap() { # tag, kernel regex, bench args...
  local tag=$1 re=$2; shift 2
  $B "$@" > gpurun_out/r2_plain_$tag.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$re -s 3 -c 2 -f -o gpurun_out/r2_prof_$tag $B "$@" > gpurun_out/r2_ncu_$tag.log 2>&1
  echo "$tag rc=$?"
}
cap sym_solve3 tile_kernel --workload sym_solve3
cap sym_solve6 tile_kernel --workload sym_solve6
cap sym_invert6 tile_kernel --workload sym_invert6
cap sym_solve10 tile_kernel --workload sym_solve10
cap dense_inv4_f64 tile_kernel --workload dense_inv4_f64
cap dense_inv8_f64 pool_kernel --kind batch_inv --n 8 --dtype f64 --batch 4194304
cap dense_inv10_f64 pool_kernel --kind batch_inv --n 10 --dtype f64 --batch 4194304
cap dense_inv10_f32 pool_kernel --kind batch_inv --n 10 --dtype f32 --batch 8388608
ls -la gpurun_out/r2_prof_*.ncu-rep | awk '{print $5, $9}'
