"""Summarise an `ncu --set full` report: one row per captured launch (markdown) + DRAM bytes per launch (json).

    python tools/ncu_summary.py gpurun_out/f_kernels.ncu-rep profiles/r01d_ncu_full_summary.md profiles/ncu_summary.json
"""
import csv
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamicvectorquantization_b200.build import source_sha  # noqa: E402

rep, out_md, out_json = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, scale=1.0):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return None
    v = float(r[i].replace(",", ""))
    u = units[i]
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "msecond": 1e3,
            "usecond": 1.0, "nsecond": 1e-3}.get(u, 1.0)
    return v * mult * scale


lines = ["| kernel | grid | time us | SM clock GHz | tensor pipe % of elapsed | DRAM read MB | DRAM write MB | DRAM GB/s | "
         "L2 hit % | L2->SM GB | issue active % | regs |", "|---|---|---|---|---|---|---|---|---|---|---|---|"]
summary = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "").replace("b2::", "")
    t = val(r, "gpu__time_duration.sum")
    cyc = val(r, "sm__cycles_elapsed.max") or val(r, "sm__cycles_active.max")
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    l2sm = val(r, "l1tex__m_xbar2l1tex_read_bytes.sum")
    lines.append("| `%s` | %s | %.1f | %s | %.1f | %.1f | %.1f | %.1f | %.1f | %.2f | %.1f | %d |" % (
        short, r[col["Grid Size"]], t, ("%.2f" % (cyc / t / 1e3)) if cyc else "-",
        val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed") or 0.0, rd / 1e6, wr / 1e6,
        (rd + wr) / t / 1e3, val(r, "lts__t_sector_hit_rate.pct") or 0.0,
        (l2sm or 0.0) / 1e9, val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") or 0.0,
        int(val(r, "launch__registers_per_thread") or 0)))
    key = ("pconv" if ("pconv" in short and ("<3>" in short or "<" not in short)) else "pconv_taps" if "pconv" in short
           else "vq" if "vq_search" in short else "wgrad" if "mmgemm" in short else "gn" if "gn_" in short else "tapgemm")
    if key == "gn":                                   # the fused GroupNorm backward is the family's dominant kernel
        if "bwd_fused" not in short:
            continue
    summary[key + "_dram_bytes_per_launch"] = int(rd + wr)
    summary[key + "_time_us"] = t
    summary[key + "_csrc_sha16"] = source_sha(key)   # run from the tree the capture was built from
if os.path.exists(out_json):                          # keep the entries of kernels this report does not contain
    old = json.load(open(out_json))
    for k, v in old.items():
        summary.setdefault(k, v)
open(out_md, "w").write("\n".join(lines) + "\n")
summary["source"] = rep.split("/")[-1] + " (dram__bytes_read.sum + dram__bytes_write.sum per launch)"
summary["note"] = ("*_csrc_sha16 = sha256[:16] of the kernel's sources (its .cu + common.cuh + tmap.h) at capture time; "
                   "bench.py reports a traffic figure only when it matches the sources it runs")
json.dump(summary, open(out_json, "w"), indent=1)
print("\n".join(lines))
