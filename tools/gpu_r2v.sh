#!/bin/bash
# GroupNorm backward with software-pipelined phases: parity tests in both modes, microbench A/B, step A/B
mkdir -p gpurun_out
for v in 1 0; do
  echo "== B2DQ_GN_PIPE=$v tests"
  B2DQ_GN_PIPE=$v timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_oplevel.py -x -q -m gpu -k "groupnorm or gn_" 2>&1 | tail -3
done
for v in 0 1; do
  echo "== B2DQ_GN_PIPE=$v gnf"
  B2DQ_GN_PIPE=$v timeout 200 python tools/kernel_bench.py gnf 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['hw'], d['c'], d['plan'], 'fused', d['fused_ms'], d['fused_GBs'], 'add', d['fused_add_ms'], 'split', d['split_ms'])
"
done
for v in 0 1; do
  B2DQ_GN_PIPE=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-real-loss 2>gpurun_out/r2v_bench.err | python -c "
import sys, json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('gn_pipe', $v, 'step', d['ms_per_step'], d['value'], 'gn', d['roofline_gn']['ms_per_launch'], d['roofline_gn']['frac'])"
done
