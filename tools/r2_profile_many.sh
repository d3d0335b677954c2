#!/bin/bash
mkdir -p gpurun_out
cap() { # tag n k dtype
  python tools/nrhs_one.py $2 $3 $4 > gpurun_out/r2_plain_many_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:solve_many_staged -s 2 -c 1 -f -o gpurun_out/r2_prof_many_$1 python tools/nrhs_one.py $2 $3 $4 > gpurun_out/r2_ncu_many_$1.log 2>&1
  echo "$1 rc=$?"
  python profiles/summarize_ncu.py full gpurun_out/r2_prof_many_$1.ncu-rep > gpurun_out/r2_lmdiv_$1_ncu_full.txt 2> gpurun_out/r2_lmdiv_$1_ncu_full.err
  rm -f gpurun_out/r2_prof_many_$1.ncu-rep
}
cap 6x6x6_f32 6 6 f32
cap 8x8x8_f32 8 8 f32
cap 6x6x6_f64 6 6 f64
