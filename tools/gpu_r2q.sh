#!/bin/bash
# 2 GPUs: multi-GPU parity tests + scaling bench (whole-step graph with bucketed exchange) + 1-GPU line on the same box
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
run() { # tag, extra args
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 20 --warmup 5 $2 > gpurun_out/r2q_$1.json 2> gpurun_out/r2q_$1.err
  echo "$1 rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2q_$1.json").read().strip().splitlines()[-1])
    print("$1", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("exchange","")[:200])
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/r2q_$1.err").read()[-1500:])
PY
}
run overlap ""
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-real-loss > gpurun_out/r2q_n1.json 2> gpurun_out/r2q_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2q_n1.json').read().strip().splitlines()[-1]); print('n1', d['value'], d['ms_per_step'])"
