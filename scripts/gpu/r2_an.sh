#!/bin/bash
# round 2, GPU call AN (final kernels): ncu --set full of the step's dominant convolution launches (after the fused pairs / third epilogue group)
set -u
mkdir -p gpurun_out
timeout 300 python scripts/exp/ncu_shapes.py > gpurun_out/r2an_conv_shapes.json 2> gpurun_out/r2an_conv_shapes.err && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'sweep|wgrad_stack' -c 40 -o /tmp/r2an_conv -f python scripts/exp/ncu_shapes.py > gpurun_out/r2an_conv_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r2an_conv.ncu-rep --page raw --csv > gpurun_out/r2an_conv_ncu_full.csv 2>> gpurun_out/r2an_conv_ncu.log
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2an_conv_ncu_full.csv")))
keep = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "launch__registers_per_thread", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
idx = [i for i, h in enumerate(rows[0]) if h in keep or "tensor" in h]
csv.writer(open("gpurun_out/r2an_conv_ncu.csv", "w")).writerows([[r[i] for i in idx] for r in rows])
PY
rm -f gpurun_out/r2an_conv_ncu_full.csv
# warp-state / stall summary of the fused pair launch (128+160), for the notes
ncu -i /tmp/r2an_conv.ncu-rep --page details --csv 2>/dev/null | grep -i "pair_sweep" | grep -i "stall\|issue\|Warp Cycles\|Eligible\|No Eligible\|Executed Ipc\|Registers\|Theoretical Occ" | head -60 > gpurun_out/r2an_pair_details.csv
wc -l gpurun_out/r2an_pair_details.csv
du -sh gpurun_out
