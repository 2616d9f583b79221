#!/bin/bash
# Evidence run: ncu --set full of the headline kernels, launch lists (surrogate + real loss), whole suite + bench.
# usage: bash tools/gpu_evidence.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pconv|tapgemm|mmgemm|vq_search|gn_bwd_fused|gn_fwd_fused" -s 11 -c 11 -o gpurun_out/${tag}_kernels -f python tools/ncu_kernels.py > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
bash tools/gpu_suite.sh ${tag}
bash tools/launch_list.sh real > gpurun_out/${tag}_launch_real.txt 2>&1
cp gpurun_out/step_launches_by_kernel.csv gpurun_out/${tag}_real_step_launches_by_kernel.csv
cp gpurun_out/step_launches_raw.csv gpurun_out/${tag}_real_step_launches_raw.csv
tail -1 gpurun_out/${tag}_launch_real.txt
timeout 300 python bench.py --aux > gpurun_out/${tag}_aux.json 2> gpurun_out/${tag}_aux.err; wc -l gpurun_out/${tag}_aux.json
