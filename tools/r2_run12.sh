#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest12.log
tail -5 gpurun_out/r2_pytest12.log
python tools/fused_timing.py > gpurun_out/r2_fused_timing.log 2>&1; cat gpurun_out/r2_fused_timing.log
python tools/unaligned_timing.py > gpurun_out/r2_unaligned_timing.log 2>&1; cat gpurun_out/r2_unaligned_timing.log
python tools/fused_update_timing.py > gpurun_out/r2_fused_update_timing.log 2>&1; cat gpurun_out/r2_fused_update_timing.log
