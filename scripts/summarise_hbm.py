#!/usr/bin/env python
"""profiles/r2_ncu_hbm.txt from gpurun_out/<tag>_hbm.json (CUDA-event timing at batch 64, scripts/prof_hbm.py) and
gpurun_out/<tag>_hbm_ncu.csv (`ncu --set full` of the same script at batch 16, exported with --page raw --csv)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2a"
ncu_tag = sys.argv[2] if len(sys.argv) > 2 else tag      # the ncu capture may come from an earlier call than the event timing
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
out = []
out.append("HBM-bound kernels of the step, round 2 (%s).  Peak = measured copy bandwidth %.0f GB/s (MEASURED_PEAKS.json)." % (tag, peak))
out.append("")
out.append("A. CUDA events, batch 64 (metrics: 8 x 3 x 512^2), 10 back-to-back calls after a warm-up; algorithmic bytes = every input")
out.append("   element read once + every output element written once (SURVEY 8d).  python scripts/prof_hbm.py")
out.append("")
out.append("%-66s %9s %9s %6s %s" % ("op", "ms", "GB/s", "frac", "launches"))
for line in open(os.path.join(ROOT, "gpurun_out", tag + "_hbm.json")):
    d = json.loads(line)
    out.append("%-66s %9.3f %9.0f %6.2f %d" % (d["op"][:66], d["ms"], d["gbs"], d["frac_of_copy_bw"], d["launches"]))
path = os.path.join(ROOT, "gpurun_out", ncu_tag + "_hbm_ncu.csv")
if os.path.isfile(path):
    rows = list(csv.reader(open(path)))
    hdr = {h: i for i, h in enumerate(rows[0])}
    out.append("")
    out.append("B. ncu --set full --clock-control none (call %s), the same script with --once --n 16 (second launch of each kernel shown); dram" % ncu_tag)
    out.append("   GB/s = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration.sum; cold-cache, serialised launches.")
    out.append("")
    out.append("%-40s %14s %9s %9s %9s %9s %7s %7s" % ("kernel", "grid", "us", "rd MB", "wr MB", "dram GB/s", "of peak", "warps%"))
    seen = {}
    for r in rows[2:]:
        name = r[hdr["Kernel Name"]].split("(")[0].replace("void ", "")
        key = (name, r[hdr["Grid Size"]])
        seen[key] = r
    for (name, grid), r in seen.items():
        rd, wr, us = float(r[hdr["dram__bytes_read.sum"]]), float(r[hdr["dram__bytes_write.sum"]]), float(r[hdr["gpu__time_duration.sum"]])
        gbs = (rd + wr) / us * 1e3
        out.append("%-40s %14s %9.1f %9.1f %9.1f %9.0f %7.2f %7.1f" % (name[:40], grid, us, rd, wr, gbs * 1e0, gbs / peak,
                                                                       float(r[hdr["sm__warps_active.avg.pct_of_peak_sustained_active"]])))
open(os.path.join(ROOT, "profiles", "r2_ncu_hbm.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
