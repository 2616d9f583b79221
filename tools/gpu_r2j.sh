#!/bin/bash
# round 2, call J: cluster-multicast strip weight gradient - parity, A/B timing, step bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv or wgrad or bias_grad or upsample or tapgemm" > gpurun_out/r2j_tests.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_tests.txt; tail -4 gpurun_out/r2j_tests.txt
cat > /tmp/wg.py <<'PY'
import os, sys, json, torch
sys.path.insert(0, os.getcwd())
from dynamicvectorquantization_b200 import kernels as kn
BF = torch.bfloat16
for nb, h, w, cin, cout in ((32, 256, 256, 128, 128), (32, 128, 128, 128, 128), (32, 128, 128, 256, 128), (32, 64, 64, 256, 256)):
    x = torch.randn(nb, h, w, cin, device="cuda").to(BF); dy = torch.randn(nb, h, w, cout, device="cuda").to(BF)
    fn = lambda: kn.conv_wgrad(x, dy, 3, 1, want_bias=True)
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(10): fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(json.dumps(dict(k="wgrad3x3", cluster=os.environ.get("B2DQ_WGRAD_CLUSTER", "1"), shape=[nb, h, w, cin, cout], ms=round(ms, 4),
                          tflops=round(2.0 * nb * h * w * cin * cout * 9 / ms / 1e9, 1))), flush=True)
PY
B2DQ_WGRAD_CLUSTER=0 timeout 120 python /tmp/wg.py > gpurun_out/r2j_wgrad.txt 2>&1

cat gpurun_out/r2j_wgrad.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-real-loss > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
head -c 330 gpurun_out/r2j_bench.json; echo; tail -2 gpurun_out/r2j_bench.err
timeout 300 python tools/kernel_bench.py conv > gpurun_out/r2j_conv.txt 2>&1; cat gpurun_out/r2j_conv.txt | cut -c1-330
