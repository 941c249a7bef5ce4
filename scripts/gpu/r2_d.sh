#!/bin/bash
# round 2, GPU call D: parity suite (1x1 tcgen05, batched eval, alternating order), order experiment, benches, conv ncu
set -u
mkdir -p gpurun_out
rm -f gpurun_out/reference_callers.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -8 gpurun_out/r2d_pytest.log
timeout 300 python scripts/exp/rdb_order.py 64 > gpurun_out/r2d_rdb_order.txt 2>&1; echo "order rc=$?"; cat gpurun_out/r2d_rdb_order.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
SRCGAN_B200_NO_ALT_ORDER=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2d_bench_noalt.json 2> gpurun_out/r2d_bench_noalt.err; echo "bench noalt rc=$?"
for wl in cascade eval; do
  timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 > gpurun_out/r2d_bench_$wl.json 2> gpurun_out/r2d_bench_$wl.err; echo "bench $wl rc=$?"
done
timeout 300 python scripts/exp/ncu_shapes.py > gpurun_out/r2d_conv_shapes.json 2> gpurun_out/r2d_conv_shapes.err && \
timeout 1200 ncu --set full --clock-control none -k regex:'sweep2|wgrad_stack' -c 40 -o /tmp/r2d_conv -f python scripts/exp/ncu_shapes.py > gpurun_out/r2d_conv_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r2d_conv.ncu-rep --page raw --csv > gpurun_out/r2d_conv_ncu_full.csv 2>> gpurun_out/r2d_conv_ncu.log
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2d_conv_ncu_full.csv")))
keep = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "launch__registers_per_thread", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
idx = [i for i, h in enumerate(rows[0]) if h in keep or "tensor" in h]
csv.writer(open("gpurun_out/r2d_conv_ncu.csv", "w")).writerows([[r[i] for i in idx] for r in rows])
PY
rm -f gpurun_out/r2d_conv_ncu_full.csv
du -sh gpurun_out
