#!/bin/bash
# ncu launch list (gpu__time_duration only) of ONE training step -> gpurun_out/step_launches.csv
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/step_launches_ncu.csv python tools/profile_step.py 32 $1 2>&1 | tail -2
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/step_launches_ncu.csv") if l.startswith('"')))
hdr = rows[0]
ik, ig, ib, iv = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size"), hdr.index("Metric Value")
iu = hdr.index("Metric Unit")
out = [("kernel", "grid", "block", "gpu_time_ns")]
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(u, 1)
    out.append((r[ik], r[ig], r[ib], int(ns)))
    a = agg.setdefault(r[ik], [0, 0]); a[0] += 1; a[1] += ns
csv.writer(open("gpurun_out/step_launches_raw.csv", "w")).writerows(out)
tot = sum(a[1] for a in agg.values())
w = csv.writer(open("gpurun_out/step_launches_by_kernel.csv", "w"))
w.writerow(("kernel", "launches", "total_ms", "share_pct"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    w.writerow((k, a[0], round(a[1] / 1e6, 4), round(100 * a[1] / tot, 2)))
print("launches", len(out) - 1, "summed ms", tot / 1e6)
PY
head -30 gpurun_out/step_launches_by_kernel.csv | cut -c1-150
