#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest11.log
tail -4 gpurun_out/r2_pytest11.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench11_default.json 2> gpurun_out/r2_bench11_default.err; tail -c 600 gpurun_out/r2_bench11_default.err
python bench.py --workload dense_inv4_f64 --steps 10 --warmup 3 > gpurun_out/r2_bench11_inv4.json 2> gpurun_out/r2_bench11_inv4.err; tail -c 600 gpurun_out/r2_bench11_inv4.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench11_ref.json 2> gpurun_out/r2_bench11_ref.err; tail -c 600 gpurun_out/r2_bench11_ref.err
NG=1 bash tools/r2_multigpu.sh
