#!/bin/bash
# round 2, call E: pconv staged epilogue - parity + per-kernel timing + step bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "pconv or conv_fwd or groupnorm" > gpurun_out/r2e_tests.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_tests.txt
tail -5 gpurun_out/r2e_tests.txt
timeout 300 python tools/kernel_bench.py pconv > gpurun_out/r2e_pconv.txt 2>&1; cat gpurun_out/r2e_pconv.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-real-loss > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
head -c 400 gpurun_out/r2e_bench.json; echo
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_modules.py -m gpu -q -x > gpurun_out/r2e_model_tests.txt 2>&1
tail -3 gpurun_out/r2e_model_tests.txt
