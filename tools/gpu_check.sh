#!/bin/bash
# One GPU box visit: kernel tests, per-kernel timing, an ncu capture of the VQ kernel, the whole GPU suite.
mkdir -p gpurun_out
echo "== vq tests" ; timeout 420 python -m pytest tests/test_gpu_kernels.py -x -q -k "vq" 2>&1 | tail -5 | tee gpurun_out/b_vq_tests.txt
echo "== kernel bench"; timeout 200 python tools/kernel_bench.py vq 2>&1 | tee gpurun_out/b_vq_bench_v2.txt
echo "== ncu vq"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:vq_search -s 2 -c 1 -o gpurun_out/b_vq_v2 -f python tools/ncu_kernels.py vq 2>&1 | tail -3
echo "== gpu suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/b_gpu_suite.txt
