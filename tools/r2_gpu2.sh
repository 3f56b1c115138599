#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m "gpu and slow" -x -q > gpurun_out/r2_pytest_slow.log 2>&1; echo "slow pytest rc=$?" | tee -a gpurun_out/r2_pytest_slow.log
tail -3 gpurun_out/r2_pytest_slow.log
python -m pytest tests/test_gpu_parity.py tests/test_gpu_team.py -m "gpu and not slow" -x -q -k "alpha or team" > gpurun_out/r2_pytest2.log 2>&1; echo "pytest2 rc=$?" | tee -a gpurun_out/r2_pytest2.log
tail -3 gpurun_out/r2_pytest2.log
bash tools/sanitize.sh gpurun_out
