#!/bin/bash
mkdir -p gpurun_out
echo "== loss tests"; timeout 900 python -m pytest tests/test_gpu_loss.py -x -q 2>&1 | tail -12 | tee gpurun_out/d_loss_tests.txt
bash tools/launch_list.sh real
mv gpurun_out/step_launches_by_kernel.csv gpurun_out/real_step_launches_by_kernel.csv
mv gpurun_out/step_launches_raw.csv gpurun_out/real_step_launches_raw.csv
