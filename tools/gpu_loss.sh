#!/bin/bash
mkdir -p gpurun_out
echo "== loss tests"; timeout 900 python -m pytest tests/test_gpu_loss.py -x -q 2>&1 | tail -12 | tee gpurun_out/d_loss_tests.txt
echo "== bench real loss"; timeout 600 python bench.py --loss real --steps 5 --warmup 3 2>gpurun_out/d_bench_real.err | tee gpurun_out/d_bench_real.json; tail -3 gpurun_out/d_bench_real.err
