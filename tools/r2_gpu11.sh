#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m "gpu and not slow" -x -q > gpurun_out/r2_pytest11.log 2>&1; echo "pytest11 rc=$?" | tee -a gpurun_out/r2_pytest11.log
tail -5 gpurun_out/r2_pytest11.log
b() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; }
b r2g_c1 python bench.py --workload c1 --no-cpu-baseline --no-parity-check
b r2g_c2 python bench.py --workload c2 --no-cpu-baseline --no-parity-check
b r2g_c4 python bench.py --workload c4 --no-cpu-baseline --no-parity-check
for f in gpurun_out/r2g_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d.get("kernels",{})
    print(sys.argv[1].split('/')[-1], "ms/step %.4f burst %.4f value %.1f e2e %.1f R=%s launches/solve %.0f" % (d["ms_per_step"], d["burst"]["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["repeats"], d["gpu_launches_per_solve"]),
          {n:(round(v["avg_ms"],4), round(v["achieved_gbs"])) for n,v in k.items()}, "reorth", d["reorth_steps"])
except Exception as e:
    print(sys.argv[1], "unreadable", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
# ncu: K5 at 256^3 (one launch), full set
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ritz_lift_gemm -c 1 -o gpurun_out/r2_k5 -f python tools/k5_time.py 256 60 > gpurun_out/r2_ncu_k5.log 2>&1; echo "ncu k5 rc=$?"
