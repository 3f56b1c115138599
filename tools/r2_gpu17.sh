#!/bin/bash
# windowed SELL on row shards: team tests, then config 4 at all the GPUs of the box
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 500 python -m pytest tests/test_gpu_team.py -x -q -m "gpu and not slow" -k "sparse or windowed or dropin" > gpurun_out/r2w_team.log 2>&1; echo "team pytest rc=$?"; tail -2 gpurun_out/r2w_team.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload c4 --no-cpu-baseline > gpurun_out/r2w_c4_n$N.json 2> gpurun_out/r2w_c4_n$N.err; echo "c4 n$N rc=$?"
LZ_SELL_WINDOW=0 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --workload c4 --no-cpu-baseline --no-parity-check > gpurun_out/r2w_c4_n${N}_plain.json 2> gpurun_out/r2w_c4_n${N}_plain.err; echo "c4 plain n$N rc=$?"
python - <<PY
import json
for f in ("r2w_c4_n$N", "r2w_c4_n${N}_plain"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], {k: (round(v["avg_ms"], 4)) for k, v in d.get("kernels", {}).items()},
              "win", d["config"].get("windowed_spmv_granules"), "overlap", d.get("overlap"), "parity", d.get("parity", {}).get("ok"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -2 gpurun_out/r2w_c4_n$N.err
