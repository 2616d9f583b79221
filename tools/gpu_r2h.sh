#!/bin/bash
# round 2, call H: fused GroupNorm forward + fused q/k/v attention node - parity, timing, step bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_modules.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/r2h_tests.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_tests.txt
tail -6 gpurun_out/r2h_tests.txt
timeout 300 python tools/kernel_bench.py gnf > gpurun_out/r2h_gnf.txt 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2h_gnf.txt'):
    if l.startswith('{'):
        d=json.loads(l); print(d['hw'],d['c'],'bwd fused',d['fused_ms'],'split',d['split_ms'],'| fwd fused',d.get('fwd_fused_ms'),d.get('fwd_fused_GBs'),'split',d.get('fwd_split_ms'))
PY
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-real-loss > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
head -c 330 gpurun_out/r2h_bench.json; echo; tail -2 gpurun_out/r2h_bench.err
