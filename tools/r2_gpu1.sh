#!/bin/bash
# round-2 GPU call 1: parity suite, then A/B of the new step (PDL, alpha inside KB) on config 3
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
python -m pytest tests -m "gpu and not slow" -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
b() { name=$1; shift; "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; }
b r2a_c3_k20 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
b r2a_c3_k20_noalpha python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kb-alpha --no-parity-check
LZ_PDL=0 b r2a_c3_k20_noalpha_nopdl python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kb-alpha --no-parity-check
b r2a_c3_k100 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-parity-check
b r2a_c3_k100_noalpha python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-kb-alpha --no-parity-check
b r2a_c1 python bench.py --workload c1 --no-cpu-baseline --no-parity-check
for f in gpurun_out/r2a_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d.get("kernels",{})
    print(sys.argv[1].split('/')[-1], "ms/step %.4f burst %.4f value %.1f e2e %.1f R=%s launches/solve %.0f" % (d["ms_per_step"], d["burst"]["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["repeats"], d["gpu_launches_per_solve"]),
          {n:(round(v["avg_ms"],4), round(v["achieved_gbs"])) for n,v in k.items()}, d["clocks"], "parity", (d.get("parity_check") or {}).get("ok"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
