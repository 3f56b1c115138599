#!/bin/bash
# config 4 at all the GPUs of the box (windowed SELL, bank-aware order)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload c4 --no-cpu-baseline > gpurun_out/r2x_c4_n$N.json 2> gpurun_out/r2x_c4_n$N.err; echo "c4 n$N rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2x_c4_n$N.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "burst", d["burst"]["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], {k: round(v["avg_ms"], 4) for k, v in d.get("kernels", {}).items()}, "parity", d["parity_check"]["ok"])
PY
