#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m "gpu and not slow" -x -q > gpurun_out/r2_pytest14.log 2>&1; echo "pytest14 rc=$?" | tee -a gpurun_out/r2_pytest14.log
tail -4 gpurun_out/r2_pytest14.log
b() { name=$1; shift; timeout 400 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; }
b r2j_c4 python bench.py --workload c4 --no-cpu-baseline --no-parity-check
LZ_SELL_UNIFORM=0 b r2j_c4_values python bench.py --workload c4 --no-cpu-baseline --no-parity-check
b r2j_c2 python bench.py --workload c2 --no-cpu-baseline --no-parity-check
for f in gpurun_out/r2j_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d.get("kernels",{})
    print(sys.argv[1].split('/')[-1], "ms/step %.4f burst %.4f value %.1f" % (d["ms_per_step"], d["burst"]["ms_per_step"], d["value"]),
          {n:(round(v["avg_ms"],4), round(v["achieved_gbs"])) for n,v in k.items()}, "value_free", d["config"].get("value_free_spmv"), "reorth", d["reorth_steps"], "roofline", round(d["roofline"]["frac"],3), d["roofline"]["kernel"])
except Exception as e:
    print(sys.argv[1], "unreadable", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
