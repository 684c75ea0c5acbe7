"""Summarise gpurun_out/ ncu artefacts into small tracked files under profiles/ (gpurun_out/ is scratch).

    python tools/summarize_profiles.py r01
"""
import collections
import csv
import json
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs("profiles", exist_ok=True)

# ---- launch list: per-kernel totals and shares ----
p = f"gpurun_out/launches_{tag}.csv"
if os.path.exists(p):
    rows = list(csv.reader(open(p)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        name = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(f"profiles/launches_{tag}_summary.csv", "w") as f:
        f.write("kernel,launches,total_us,share_pct,avg_us\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{v[0]},{v[1]:.1f},{100 * v[1] / tot:.2f},{v[1] / v[0]:.2f}\n")
    print(f"wrote profiles/launches_{tag}_summary.csv ({sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms)")

# ---- full capture: the metrics the roofline quotes ----
import glob
for rep in sorted(glob.glob(f"gpurun_out/prof_{tag}*.ncu-rep")):
    suffix = os.path.basename(rep)[len(f"prof_{tag}"):-len(".ncu-rep")]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
            "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__inst_executed.sum"]
    idx = [(k, hdr.index(k)) for k in keys if k in hdr]
    res = []
    for r in rows[2:]:
        res.append({k: (r[i] + (" " + units[i] if units[i] else "")) for k, i in idx})
    json.dump(res, open(f"profiles/ncu_full_{tag}{suffix}_summary.json", "w"), indent=1)
    print(f"wrote profiles/ncu_full_{tag}{suffix}_summary.json ({len(res)} kernels)")
