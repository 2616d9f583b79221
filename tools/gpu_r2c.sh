#!/bin/bash
# round 2, call C: PatchGAN on the kernels (parity, real-loss bench, launch list) + GroupNorm L2 hints
mkdir -p gpurun_out

timeout 900 python -m pytest tests/test_gpu_loss.py -m gpu -q -s > gpurun_out/r2c_loss_tests.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_loss_tests.txt
grep -E "patchgan|passed|failed|Error|assert" gpurun_out/r2c_loss_tests.txt | tail -25
timeout 600 python bench.py --loss real --steps 10 --warmup 3 > gpurun_out/r2c_real.json 2> gpurun_out/r2c_real.err
head -c 900 gpurun_out/r2c_real.json; echo; tail -3 gpurun_out/r2c_real.err
bash tools/launch_list.sh real > gpurun_out/r2c_launch_real.txt 2>&1
cp gpurun_out/step_launches_by_kernel.csv gpurun_out/r2c_real_step_launches_by_kernel.csv
cp gpurun_out/step_launches_raw.csv gpurun_out/r2c_real_step_launches_raw.csv
tail -3 gpurun_out/r2c_launch_real.txt
grep -i -E "cudnn|cutlass|implicit|batch_norm|bn_" gpurun_out/r2c_real_step_launches_by_kernel.csv | cut -c1-160
