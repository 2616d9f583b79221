#!/bin/bash
# 2 GPUs: where does the exposed exchange time come from?
mkdir -p gpurun_out
run() { # tag, env, extra args
  env $2 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 20 --warmup 5 $3 > gpurun_out/r2r_$1.json 2> gpurun_out/r2r_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2r_$1.json").read().strip().splitlines()[-1])
    print("$1", round(d["value"],1), round(d["ms_per_step"],2))
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/r2r_$1.err").read()[-800:])
PY
}
run default "A=1" ""
run nooverlap "A=1" "--no-overlap"
run ch2 "NCCL_MAX_NCHANNELS=2" ""
run ch8 "NCCL_MAX_NCHANNELS=8" ""
run onebucket "B2DQ_BUCKET_MB=400" ""
run ch4_nooverlap "NCCL_MAX_NCHANNELS=4" "--no-overlap"
