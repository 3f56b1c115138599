#!/bin/bash
mkdir -p gpurun_out
tr() { name=$1; shift; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; }
tr r2k_c4_n8 --workload c4 --steps 100
for f in gpurun_out/r2k_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d.get("kernels",{})
    print(sys.argv[1].split('/')[-1], "ms/step %.4f burst %.4f value %.1f e2e %.1f R=%s" % (d["ms_per_step"], d["burst"]["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["repeats"]),
          {n:(round(v["avg_ms"],4), round(v["achieved_gbs"])) for n,v in k.items()}, "overlap", d["fused_step"].get("overlap"), "reorth", d["reorth_steps"], "parity", (d.get("parity_check") or {}).get("ok"), d["config"].get("value_free_spmv"))
except Exception as e:
    print(sys.argv[1], "unreadable", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
