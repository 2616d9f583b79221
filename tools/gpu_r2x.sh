#!/bin/bash
# 4 GPUs: scaling bench line as the driver launches it
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 \
  bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2w_n8.json 2> gpurun_out/r2w_n8.err
echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2w_n8.json").read().strip().splitlines()[-1])
    print("n8", round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), d.get("exchange","")[:120])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2w_n8.err").read()[-1500:])
PY
