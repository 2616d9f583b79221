#!/bin/bash
# Whole GPU parity suite + headline bench (with cpu_baseline and real_loss_step) + launch list of one step.
# usage: bash tools/gpu_suite.sh <tag>
tag=${1:-suite}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${tag}_gpu_suite.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_gpu_suite.txt
tail -14 gpurun_out/${tag}_gpu_suite.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.txt 2>&1; tail -1 gpurun_out/${tag}_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print("img/s", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print("roofline", d["roofline"]["ms_per_launch"], d["roofline"]["frac"], "gn", d["roofline_gn"]["ms_per_launch"], d["roofline_gn"]["frac"])
print("vq", d["roofline_vq"]["ms_per_launch"], d["roofline_vq"]["tensor_frac"], d["roofline_vq"].get("training_shape"))
print("cpu", d.get("cpu_baseline"))
print("real", (d.get("real_loss_step") or {}).get("value"), (d.get("real_loss_step") or {}).get("ms_per_step"))
print("mfu", d["model_flops_utilisation"])
PY
bash tools/launch_list.sh > gpurun_out/${tag}_launch.txt 2>&1
cp gpurun_out/step_launches_by_kernel.csv gpurun_out/${tag}_step_launches_by_kernel.csv
cp gpurun_out/step_launches_raw.csv gpurun_out/${tag}_step_launches_raw.csv
tail -1 gpurun_out/${tag}_launch.txt
head -16 gpurun_out/${tag}_step_launches_by_kernel.csv | cut -c1-120
