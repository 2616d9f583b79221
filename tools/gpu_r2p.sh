#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for tpc in 1 2; do
B2DQ_UPCONV_WGRAD_TPC=$tpc timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-real-loss 2>gpurun_out/r2p_bench.err | python -c "
import sys, json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tpc', $tpc, 'step', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])"
done
B2DQ_UPCONV_WGRAD_TPC=2 timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "upconv or upsample" 2>&1 | tail -2
