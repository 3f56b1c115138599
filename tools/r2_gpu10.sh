#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m "gpu and not slow" -x -q -k "kba or recompute or alpha or large_grid or full_size_config3 or graph" > gpurun_out/r2_pytest10.log 2>&1; echo "pytest10 rc=$?" | tee -a gpurun_out/r2_pytest10.log
tail -5 gpurun_out/r2_pytest10.log
b() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; }
b r2f_c3_k100 python bench.py --steps 100 --no-cpu-baseline --no-parity-check
LZ_KBA=0 b r2f_c3_k100_nokba python bench.py --steps 100 --no-cpu-baseline --no-parity-check
LZ_KBA_ZC=4 b r2f_c3_k100_zc4 python bench.py --steps 100 --no-cpu-baseline --no-parity-check
LZ_KBA_ZC=16 b r2f_c3_k100_zc16 python bench.py --steps 100 --no-cpu-baseline --no-parity-check
LANCZOS_B200_LIB=$PWD/lanczos_b200/variant_kba3.so b r2f_c3_k100_mb3 python bench.py --steps 100 --no-cpu-baseline --no-parity-check
b r2f_c3_k20 python bench.py --steps 20 --no-cpu-baseline --no-parity-check
for f in gpurun_out/r2f_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d.get("kernels",{})
    print(sys.argv[1].split('/')[-1], "ms/step %.4f burst %.4f value %.1f e2e %.1f R=%s launches/solve %.0f" % (d["ms_per_step"], d["burst"]["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["repeats"], d["gpu_launches_per_solve"]),
          {n:(round(v["avg_ms"],4), round(v["achieved_gbs"])) for n,v in k.items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "unreadable", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
