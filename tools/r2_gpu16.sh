#!/bin/bash
# windowed SELL form: its tests, then config 4 with and without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_team.py -x -q -m "gpu and not slow" -k "windowed or sparse or sell or rgg or irregular" > gpurun_out/r2w_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2w_pytest.log
timeout 400 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2w_c4.json 2> gpurun_out/r2w_c4.err; echo "c4 rc=$?"
LZ_SELL_WINDOW=0 timeout 400 python bench.py --workload c4 --no-cpu-baseline --no-parity-check > gpurun_out/r2w_c4_plain.json 2> gpurun_out/r2w_c4_plain.err; echo "c4 plain rc=$?"
python - <<'PY'
import json
for f in ("r2w_c4", "r2w_c4_plain"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "value", d["value"], "kernels", d.get("kernels"), "roofline", d["roofline"]["achieved"], d["roofline"]["frac"],
              "win", d["config"].get("windowed_spmv_granules"), "parity", d.get("parity", {}).get("ok"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r2w_c4.err
