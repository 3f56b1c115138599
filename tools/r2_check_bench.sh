#!/bin/bash
# the driver's commands on the final code: gpu suite, default bench line, reference arm
mkdir -p gpurun_out
python -m pytest tests -m "gpu and not slow" -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r2y_c3.json 2> gpurun_out/r2y_c3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2y_c3.json").read().strip().splitlines()[-1])
print("ms/step %.4f burst %.4f value %.1f e2e %.1f" % (d["ms_per_step"], d["burst"]["ms_per_step"], d["value"], d["e2e"]["value"]))
print("roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],4), round(d["roofline"]["achieved"]), d["roofline"]["traffic"], d["roofline"]["traffic_source"])
print("kernels", {k:(round(v["avg_ms"],4), round(v["achieved_gbs"])) for k,v in d["kernels"].items()})
print("sustained", {k:(round(v["avg_ms"],4), v["achieved_gbs"] and round(v["achieved_gbs"])) for k,v in d["kernels_sustained"].items()})
print("clocks", d["clocks"], "parity", d["parity_check"]["ok"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"])
print("fused", d["fused_step"]["frac_of_measured_peak"], d["fused_step"]["burst_frac_of_measured_peak"])
PY
