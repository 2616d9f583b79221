#!/bin/bash
# Short end-of-session check of the final tree: whole GPU suite, smoke, headline bench, real-loss bench.
mkdir -p gpurun_out
echo "== gpu suite"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/s_gpu_suite.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/s_bench.err | tee gpurun_out/s_bench.json | cut -c1-300
echo "== real"; timeout 400 python bench.py --loss real --steps 5 --warmup 3 2>gpurun_out/s_real.err | tee gpurun_out/s_real.json | cut -c1-260
