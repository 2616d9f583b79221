#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_oplevel.py -x -q -m gpu -k "parity_classes or folded_upsample or conv2d_forward or conv_fwd_dgrad" 2>&1 | tail -4
for v in 0 1; do
  echo "== PCONV_TAPS=$v"
  B2DQ_PCONV_TAPS=$v timeout 300 python - <<'PY'
import sys, os, json
sys.path.insert(0, os.getcwd()); sys.argv=["x","none"]
exec(open("tools/kernel_bench.py").read().split("if __name__")[0])
conv(32, 256, 256, 128, 128, 3, 2)
upconv(32, 128, 128, 128, 128)
PY
done
for v in 0 1; do
  B2DQ_PCONV_TAPS=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-real-loss 2>gpurun_out/r2t_bench.err | python -c "
import sys, json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pconv_taps', $v, 'step', d['ms_per_step'], d['value'])"
done
