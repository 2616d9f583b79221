#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-real-loss 2>gpurun_out/r2o_bench.err | python -c "
import sys, json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])"
timeout 300 python bench.py --loss real --steps 10 --warmup 3 2>gpurun_out/r2o_real.err | python -c "
import sys, json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('real step', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'])"
