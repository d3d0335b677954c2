#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/nrhs_timing.py > gpurun_out/r2_nrhs_timing.log 2>&1; cat gpurun_out/r2_nrhs_timing.log
