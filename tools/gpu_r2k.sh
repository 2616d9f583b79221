#!/bin/bash
mkdir -p gpurun_out
for kb in 16 64 100000; do
  echo "== WGRAD_TWO_WAVE_MIN_KB=$kb"
  B2DQ_WGRAD_TWO_WAVE_MIN_KB=$kb timeout 200 python tools/kernel_bench.py conv 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['k'], d['hw'], d['cin'], d['cout'], 'wgrad', d['wgrad_ms'], d['wgrad_tflops'])
"
done
for kb in 16 64 100000; do
  B2DQ_WGRAD_TWO_WAVE_MIN_KB=$kb timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-real-loss 2>/dev/null | python -c "
import sys, json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('kb', $kb, 'step', d['ms_per_step'], d['value'])"
done
