#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m "gpu and not slow" -x -q > gpurun_out/r2_pytest8.log 2>&1; echo "pytest8 rc=$?" | tee -a gpurun_out/r2_pytest8.log
tail -5 gpurun_out/r2_pytest8.log
python tools/k5_time.py 512 60 2>&1 | tail -1 | tee gpurun_out/r2_k5.txt
b() { name=$1; shift; "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; }
b r2d_c1 python bench.py --workload c1 --no-cpu-baseline --no-parity-check
LZ_GRAPH=0 b r2d_c1_nograph python bench.py --workload c1 --no-cpu-baseline --no-parity-check
b r2d_c2 python bench.py --workload c2 --no-cpu-baseline --no-parity-check
for f in gpurun_out/r2d_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d.get("kernels",{})
    print(sys.argv[1].split('/')[-1], "ms/step %.4f burst %.4f value %.1f e2e %.1f R=%s launches/solve %.0f" % (d["ms_per_step"], d["burst"]["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["repeats"], d["gpu_launches_per_solve"]),
          {n:(round(v["avg_ms"],4), round(v["achieved_gbs"])) for n,v in k.items()})
except Exception as e:
    print(sys.argv[1], "unreadable", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
