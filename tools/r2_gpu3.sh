#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m "gpu and not slow" -x -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest3 rc=$?" | tee -a gpurun_out/r2_pytest3.log
tail -15 gpurun_out/r2_pytest3.log
python -m pytest tests/test_gpu_parity.py -m "gpu and slow" -x -q > gpurun_out/r2_pytest_slow.log 2>&1; echo "slow pytest rc=$?" | tee -a gpurun_out/r2_pytest_slow.log
tail -5 gpurun_out/r2_pytest_slow.log
