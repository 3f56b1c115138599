#!/bin/bash
# final state check on one GPU: the GPU suite, smoke, config 4, and the driver's bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m "gpu and not slow" > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2x_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 400 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2x_c4.json 2> gpurun_out/r2x_c4.err; echo "c4 rc=$?"
timeout 900 python bench.py > gpurun_out/r2x_c3.json 2> gpurun_out/r2x_c3.err; echo "c3 rc=$?"
python - <<'PY'
import json
for f in ("r2x_c4", "r2x_c3"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "burst", d["burst"]["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"],
              {k: round(v["avg_ms"], 4) for k, v in d.get("kernels", {}).items()}, "roofline", round(d["roofline"]["achieved"]), round(d["roofline"]["frac"], 4),
              "parity", d["parity_check"]["ok"], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
