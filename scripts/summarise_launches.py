#!/usr/bin/env python
"""profiles/<name>.csv.gz (ncu --metrics gpu__time_duration.sum --clock-control none --csv of `python bench.py --steps 1 --warmup 1
--no-cpu-baseline --no-e2e`) -> per-kernel launch counts, serialised device time and share of the whole run.  The run holds three
identical steps (warm-up, timed, instrumented) plus start-up launches; ncu's per-launch times are cold-cache and serialised, so
only the SHARES are comparable with bench.py's `roofline.share_of_step`.
python scripts/summarise_launches.py profiles/r2_launches_bench.csv.gz > profiles/r2_launches_bench_summary.txt"""
import csv
import gzip
import re
import sys
from collections import defaultdict

path = sys.argv[1]
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = [r for r in csv.reader(f) if len(r) > 10]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    name = r[ik]
    m = re.search(r"((?:srcgan|tcw4|tcw3|tcw|tc)::[\w:]+(<[^(]*>)?)", name)       # ncu drops the outer namespace of templates
    key = ("srcgan::" + m.group(1)).replace("srcgan::srcgan::", "srcgan::") if m else ("ATen / other: " + re.sub(r"\(.*", "", name)[:70])
    agg[key][0] += 1
    agg[key][1] += float(r[iv].replace(",", "")) / 1e3
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if k.startswith("srcgan"))
print("%s: %d launches, %.1f ms of serialised device time; this library's kernels: %.1f %%" % (path, n, tot / 1e3, ours / tot * 100))
fam = defaultdict(float)
for k, v in agg.items():
    for f_ in ("conv3x3_wgrad_stack_tc", "conv3x3_wgrad_r32_tc", "conv3x3_pair_sweep_tc", "conv3x3_sweep2_tc", "conv_igemm_tc", "conv_wgrad_tc_kernel", "bn_", "wgrad_reduce"):
        if f_ in k:
            fam[f_] += v[1]
print("family shares of the run: " + ", ".join("%s %.1f %%" % (k, v / tot * 100) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])))
print("\n%-100s %7s %11s %7s" % ("kernel", "count", "us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-100s %7d %11.0f %6.2f%%" % (k[:100], v[0], v[1], v[1] / tot * 100))
