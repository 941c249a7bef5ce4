#!/bin/bash
# round 2, GPU call BA: ncu --set full of the FFMA thin-layer kernels (thin_in_tiled on the 7x7 stride-2 stem and the first discriminator
# layer, thin_wgrad7_kernel) - what bounds them
set -u
mkdir -p gpurun_out
timeout 120 python scripts/exp/ncu_thin.py > gpurun_out/r2ba_thin.log 2>&1; echo "plain rc=$?"
timeout 200 ncu --set full --clock-control none -k regex:'thin_in_tiled|thin_wgrad7' -c 9 -o /tmp/r2ba_thin -f python scripts/exp/ncu_thin.py >> gpurun_out/r2ba_thin.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r2ba_thin.ncu-rep --page raw --csv > /tmp/r2ba_full.csv 2>> gpurun_out/r2ba_thin.log
python - <<'PY'
import csv
rows = list(csv.reader(open("/tmp/r2ba_full.csv")))
keep = ("ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed")
idx = [i for i, h in enumerate(rows[0]) if h in keep or "pipe_fma" in h or "mem_shared" in h]
csv.writer(open("gpurun_out/r2ba_thin_ncu.csv", "w")).writerows([[r[i] for i in idx] for r in rows])
PY
wc -l gpurun_out/r2ba_thin_ncu.csv
