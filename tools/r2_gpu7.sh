#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m "gpu and not slow" -x -q > gpurun_out/r2_pytest7.log 2>&1; echo "pytest7 rc=$?" | tee -a gpurun_out/r2_pytest7.log
tail -5 gpurun_out/r2_pytest7.log
python tools/k5_time.py 512 60 2>&1 | tail -1 | tee gpurun_out/r2_k5.txt
LZ_K5_GEMM=0 python tools/k5_time.py 512 60 2>&1 | tail -1 | tee -a gpurun_out/r2_k5.txt
python tools/k5_time.py 256 120 2>&1 | tail -1 | tee -a gpurun_out/r2_k5.txt
