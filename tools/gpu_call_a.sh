#!/bin/bash
# GPU call A: validate the role-split VQ search kernel, A/B against v1, then the whole GPU suite.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
echo "== vq tests" ; timeout 420 python -m pytest tests/test_gpu_kernels.py -x -q -k "vq" 2>&1 | tail -15 | tee gpurun_out/a_vq_tests.txt
echo "== kernel bench v2"; timeout 200 python tools/kernel_bench.py vq 2>&1 | tee gpurun_out/a_vq_bench_v2.txt
echo "== kernel bench v1"; B2DQ_VQ_V1=1 timeout 200 python tools/kernel_bench.py vq 2>&1 | tee gpurun_out/a_vq_bench_v1.txt
echo "== full gpu suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/a_gpu_suite.txt
echo "== bench"; timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/a_bench.err | tee gpurun_out/a_bench.json
echo "== aux"; timeout 300 python bench.py --aux 2>gpurun_out/a_aux.err | tee gpurun_out/a_aux.json
