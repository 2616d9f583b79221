#!/bin/bash
# TMA-store epilogues of the tap GEMM / batched GEMM and the operator-level C entries: tests, A/B microbench, step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_oplevel.py tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -15
for v in 0 1; do
  echo "== TMA_OUT=$v"
  B2DQ_TAPGEMM_TMA_OUT=$v B2DQ_MMGEMM_TMA_OUT=$v timeout 300 python tools/kernel_bench.py tap 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['k'], d['hw'], d['cin'], d['cout'], ' '.join(f'{k}={v}' for k,v in d.items() if k.endswith('_ms')))
"
done
for v in 0 1; do
  B2DQ_TAPGEMM_TMA_OUT=$v B2DQ_MMGEMM_TMA_OUT=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-real-loss 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tma_out', $v, 'step', d['ms_per_step'], d['value'])"
done
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_modules.py tests/test_gpu_loss.py -x -q -m gpu 2>&1 | tail -8
