#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/glue_report.py 32 > gpurun_out/r2n_glue.txt 2>&1; tail -70 gpurun_out/r2n_glue.txt
timeout 300 python -m pytest tests/test_gpu_modules.py tests/test_gpu_kernels.py -x -q -m gpu -k "attn or Attn or attention or layout or im2col or oplevel" 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-real-loss 2>gpurun_out/r2n_bench.err | python -c "
import sys, json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'])"
