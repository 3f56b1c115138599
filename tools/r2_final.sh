#!/bin/bash
# round-2 profile set (one B200): bench lines, ncu launch list, ncu --set full captures of the top kernels
mkdir -p gpurun_out
b() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; }
b r2z_c3 python bench.py --steps 20 --warmup 5
b r2z_ref python bench.py --impl reference --steps 20 --warmup 5
b r2z_c3_k100 python bench.py --steps 100 --no-cpu-baseline --no-parity-check
b r2z_c1 python bench.py --workload c1 --no-parity-check
b r2z_c2 python bench.py --workload c2 --no-cpu-baseline --no-parity-check
b r2z_c4 python bench.py --workload c4 --no-cpu-baseline --no-parity-check
b r2z_deut python bench.py --workload deut --no-parity-check
b r2z_c3full python bench.py --workload c3full --no-cpu-baseline --no-parity-check
python tools/k5_time.py 512 60 2>&1 | tail -1 > gpurun_out/r2z_k5.txt
# launch list of the default bench command (kernel names + durations)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches_c3.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check --min-region-s 0.01 > gpurun_out/r2z_ncu_launches.log 2>&1; echo "launch list rc=$?"
# full captures
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stencil_apply_dot_kernel|stencil_alpha_fast_kernel" -s 6 -c 4 \
    -o gpurun_out/r2z_c3_kernels -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity-check --min-region-s 0.01 > gpurun_out/r2z_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spmv_sell_dot_kernel|update_norm_kernel" -s 8 -c 2 \
    -o gpurun_out/r2z_c4_kernels -f python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu-baseline --no-parity-check --min-region-s 0.01 > gpurun_out/r2z_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stencil27_apply_dot_kernel|cgs_dots_kernel|cgs_update_kernel" -s 60 -c 3 \
    -o gpurun_out/r2z_deut_kernels -f python bench.py --workload deut --steps 40 --warmup 3 --no-cpu-baseline --no-parity-check --min-region-s 0.01 > gpurun_out/r2z_ncu_deut.log 2>&1; echo "ncu deut rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ritz_lift_gemm -c 1 -o gpurun_out/r2z_k5 -f python tools/k5_time.py 256 60 > gpurun_out/r2z_ncu_k5.log 2>&1; echo "ncu k5 rc=$?"
for f in gpurun_out/r2z_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    if d.get("impl") == "reference":
        print(sys.argv[1].split('/')[-1], "reference value", d["value"], d["cpu_baseline"]["kind"]); raise SystemExit
    k=d.get("kernels",{})
    print(sys.argv[1].split('/')[-1], "ms/step %.5f burst %.5f value %.1f e2e %.1f R=%s launches/solve %.0f" % (d["ms_per_step"], d["burst"]["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["repeats"], d["gpu_launches_per_solve"]),
          {n:(round(v["avg_ms"],4), round(v["achieved_gbs"])) for n,v in k.items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "roofline", d["roofline"] and round(d["roofline"]["frac"],3), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except SystemExit:
    pass
except Exception as e:
    print(sys.argv[1], "unreadable", e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
