#!/usr/bin/env python
"""gpurun_out/<tag>_conv_ncu.csv (ncu --page raw --csv of scripts/exp/ncu_shapes.py) + gpurun_out/<tag>_conv_shapes.json ->
profiles/r2_ncu_conv.txt and profiles/traffic.json (what bench.py puts into roofline.traffic)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2d"
shapes = [json.loads(l) for l in open(os.path.join(ROOT, "gpurun_out", tag + "_conv_shapes.json")) if l.startswith("{")]
rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", tag + "_conv_ncu.csv"))))
hdr = {h: i for i, h in enumerate(rows[0])}
launches = [r for r in rows[2:] if any(k in r[hdr["Kernel Name"]] for k in ("sweep", "wgrad_stack", "wgrad_r32"))]
out_name = sys.argv[2] if len(sys.argv) > 2 else "r2_ncu_conv.txt"
out, traffic, i = [], {}, 0
UNIT = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def get(r, k):
    """value of column k; bytes in MB and times in us whatever unit ncu picked for the column"""
    if k not in hdr or r[hdr[k]] in ("", "n/a"):
        return float("nan")
    return float(r[hdr[k]]) * UNIT.get(rows[1][hdr[k]], 1.0)
out.append("ncu --set full --clock-control none, scripts/exp/ncu_shapes.py (last of %d launches per shape); B200, round 2 (%s)" % (shapes[0]["launches"], tag))
out.append("%-46s %9s %9s %9s %9s %7s %9s %8s %8s" % ("shape", "us", "rd MB", "wr MB", "alg MB", "x alg", "TFLOP/s", "pipe act%", "utchmma%"))
out.append("(pipe act% = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active; utchmma% = sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32"
           "_sparsity_off.avg.pct_of_peak_sustained_elapsed; x alg = measured DRAM bytes / algorithmic bytes)")
for s in shapes:
    grp = launches[i:i + s["launches"]]
    i += s["launches"]
    if not grp:
        continue
    r = grp[-1]
    us, rd, wr = get(r, "gpu__time_duration.sum"), get(r, "dram__bytes_read.sum"), get(r, "dram__bytes_write.sum")
    alg = s["algorithmic_bytes"] / 1e6
    tens = get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    util = get(r, "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed")
    out.append("%-46s %9.1f %9.1f %9.1f %9.1f %7.2f %9.0f %8.1f %8.1f" % (s["what"], us, rd, wr, alg, (rd + wr) / alg, s["flops"] / us / 1e6, tens, util))
    name = r[hdr["Kernel Name"]]
    fam = s["kernel"]
    # the shape bench.py names for a family: the one that carries most of the family's time in the step
    prefer = {"conv3x3_wgrad_stack_tc": "wgrad 192->64", "conv3x3_pair_sweep_tc": "128->32 + 160->32", "conv3x3_sweep2_tc<64,2>": "192->64",
              "conv3x3_sweep2_tc<32,2>": "64->32 @ 64x256"}
    if fam not in traffic or prefer.get(fam, "\0") in s["what"]:
        traffic[fam] = {"bytes_per_launch": (rd + wr) * 1e6, "note": "%s: dram read %.0f MB + write %.0f MB per launch vs %.0f MB algorithmic, tensor pipe active %.1f %% (ncu --set full, profiles/%s)" % (s["what"], rd, wr, alg, tens, out_name)}
open(os.path.join(ROOT, "profiles", out_name), "w").write("\n".join(out) + "\n")
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print("\n".join(out))
print(json.dumps(traffic, indent=1))
