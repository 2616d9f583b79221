#!/bin/bash
# Round-end evidence run: whole GPU suite, headline bench (+cpu baseline), widened rows, ncu captures, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/f_smi.txt 2>&1
echo "== gpu suite"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/f_gpu_suite.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/f_smoke.txt
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/f_bench.err | tee gpurun_out/f_bench.json
echo "== aux"; timeout 300 python bench.py --aux 2>gpurun_out/f_aux.err | tee gpurun_out/f_aux.json
echo "== real loss"; timeout 400 python bench.py --loss real --steps 10 --warmup 3 2>gpurun_out/f_real.err | tee gpurun_out/f_real.json
echo "== kernel bench"; timeout 300 python tools/kernel_bench.py vq 2>&1 | tee gpurun_out/f_kernel_bench.txt
echo "== ncu full"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pconv|tapgemm|mmgemm|vq_search" -s 4 -c 4 -o gpurun_out/f_kernels -f python tools/ncu_kernels.py 2>&1 | tail -2
echo "== launch list"; bash tools/launch_list.sh 2>&1 | tail -3
