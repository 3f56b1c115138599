#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m "gpu and not slow" -x -q > gpurun_out/r2_pytest4.log 2>&1; echo "pytest4 rc=$?" | tee -a gpurun_out/r2_pytest4.log
tail -12 gpurun_out/r2_pytest4.log
