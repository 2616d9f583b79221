#!/bin/bash
# round 2, call A: GPU parity suite (with the new audits printed) + headline bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s --durations=15 > gpurun_out/r2a_gpu_suite.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_gpu_suite.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_gpu_suite.txt
cat gpurun_out/r2a_bench.json | head -c 3000
