#!/bin/bash
# round 2, call B: fused GroupNorm backward - parity, timing per shape and L2 budget, ncu, step bench
mkdir -p gpurun_out
rm -f gpurun_out/r2b_gnf.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "groupnorm" > gpurun_out/r2b_gn_tests.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_gn_tests.txt
tail -3 gpurun_out/r2b_gn_tests.txt
for mb in 70 140; do
  B2DQ_GN_L2_BUDGET_MB=$mb timeout 300 python tools/kernel_bench.py gnf >> gpurun_out/r2b_gnf.txt 2>&1
done
cat gpurun_out/r2b_gnf.txt | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-real-loss > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
head -c 600 gpurun_out/r2b_bench.json; echo
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_modules.py -m gpu -q -x > gpurun_out/r2b_model_tests.txt 2>&1
tail -3 gpurun_out/r2b_model_tests.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_bwd_fused -s 3 -c 2 -o gpurun_out/r2b_gnf -f python tools/kernel_bench.py gnf > gpurun_out/r2b_ncu.log 2>&1
tail -2 gpurun_out/r2b_ncu.log
