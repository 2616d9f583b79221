#!/bin/bash
# One GPU box visit: the whole GPU suite, per-kernel timing, the headline bench.  Extra args: "ncu" adds
# an ncu --set full capture of the VQ kernel, "aux" the widened-row timings.
mkdir -p gpurun_out
echo "== gpu suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/c_gpu_suite.txt
echo "== kernel bench"; timeout 200 python tools/kernel_bench.py vq 2>&1 | tee gpurun_out/c_vq_bench.txt
echo "== bench"; timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/c_bench.err | tee gpurun_out/c_bench.json
for a in "$@"; do
  if [ "$a" = ncu ]; then
    echo "== ncu vq"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:vq_search -s 2 -c 1 -o gpurun_out/c_vq -f python tools/ncu_kernels.py vq 2>&1 | tail -3
  fi
  if [ "$a" = aux ]; then
    echo "== aux"; timeout 300 python bench.py --aux 2>gpurun_out/c_aux.err | tee gpurun_out/c_aux.json
  fi
done
