#!/bin/bash
# round 2, call F: VQ at K=1024 - grid / split sweep and an ncu source-level capture
mkdir -p gpurun_out
timeout 300 python tools/kernel_bench.py vqk > gpurun_out/r2f_vqk.txt 2>&1
B2DQ_VQ_SPLIT_MIN_TILES=4 timeout 300 python tools/kernel_bench.py vqk >> gpurun_out/r2f_vqk.txt 2>&1
cat gpurun_out/r2f_vqk.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vq_search -s 2 -c 1 -o gpurun_out/r2f_vq -f python tools/ncu_kernels.py vq > gpurun_out/r2f_ncu.log 2>&1
tail -2 gpurun_out/r2f_ncu.log
