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
  # gpurun brings back at most 64 MiB: keep the condensed text, and the .ncu-rep of two kernels only
  python profiles/summarize_ncu.py full gpurun_out/r2_prof_$tag.ncu-rep > gpurun_out/r2_${tag}_ncu_full.txt 2> gpurun_out/r2_${tag}_ncu_full.err
  case $tag in sym_solve3|dense_inv8_f64) ;; *) rm -f gpurun_out/r2_prof_$tag.ncu-rep ;; esac
}
cap sym_solve3 tile_kernel --workload sym_solve3
cap sym_solve6 tile_kernel --workload sym_solve6
cap sym_invert6 tile_kernel --workload sym_invert6
cap sym_solve10 tile_kernel --workload sym_solve10
cap dense_inv4_f64 tile_kernel --workload dense_inv4_f64
cap dense_inv8_f64 pool_kernel --kind batch_inv --n 8 --dtype f64 --batch 4194304
cap dense_inv10_f64 pool_kernel --kind batch_inv --n 10 --dtype f64 --batch 4194304
cap dense_inv10_f32 pool_kernel --kind batch_inv --n 10 --dtype f32 --batch 8388608
python profiles/summarize_ncu.py list gpurun_out/r2_launches_default.csv > gpurun_out/r2_sym_solve3_launches.txt 2>&1
du -sh gpurun_out; ls gpurun_out
