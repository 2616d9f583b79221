#!/bin/bash
# round 2, call G (2 GPUs): multi-GPU parity test + bench with / without overlapped gradient exchange + plain DDP
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2g_smi.txt
run() { # tag, extra args
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 20 --warmup 5 $2 > gpurun_out/r2g_$1.json 2> gpurun_out/r2g_$1.err
  echo "$1 rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2g_$1.json").read().strip().splitlines()[-1])
    print("$1", d["value"], d["ms_per_step"], d.get("exchange","")[:160])
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/r2g_$1.err").read()[-1500:])
PY
}
run overlap ""
run nooverlap "--no-overlap"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-real-loss > gpurun_out/r2g_n1.json 2> gpurun_out/r2g_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2g_n1.json').read().strip().splitlines()[-1]); print('n1', d['value'], d['ms_per_step'])"
