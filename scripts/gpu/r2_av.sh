#!/bin/bash
# round 2, GPU call AV (final build of the round): whole GPU suite + smoke, the four bench workloads, per-shape profile, ncu launch list of
# the bench command, ncu --set full of the dominant convolution launches (traffic.json)
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2av_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2av_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2av_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2av_smoke.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/r2av_bench.json 2> gpurun_out/r2av_bench.err; echo "bench rc=$?"
for wl in cascade cascade_lab eval; do
  timeout 600 python bench.py --workload $wl > gpurun_out/r2av_bench_$wl.json 2> gpurun_out/r2av_bench_$wl.err; echo "bench $wl rc=$?"
done
for f in gpurun_out/r2av_bench.json gpurun_out/r2av_bench_cascade.json gpurun_out/r2av_bench_cascade_lab.json gpurun_out/r2av_bench_eval.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d["roofline"]
print(sys.argv[1], round(d["value"],1), d["unit"], round(d["ms_per_step"],2), "ms | e2e", round(d["e2e"]["value"],1), "| clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "| roofline", r["kernel"], round(r["frac"],3), "| cpu", round(d["cpu_baseline"]["value"],3), d["cpu_baseline"]["kind"])
PY
done
timeout 300 python scripts/profile_shapes.py > gpurun_out/r2av_profile_shapes.txt 2> gpurun_out/r2av_profile_shapes.err; echo "shapes rc=$?"
timeout 300 python scripts/profile_step.py > gpurun_out/r2av_profile_step.txt 2> gpurun_out/r2av_profile_step.err; echo "profile_step rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2av_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2av_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
gzip -f gpurun_out/r2av_launches.csv
timeout 300 python scripts/exp/ncu_shapes.py > gpurun_out/r2av_conv_shapes.json 2> gpurun_out/r2av_conv_shapes.err && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'sweep|wgrad_stack|wgrad_r32' -c 40 -o /tmp/r2av_conv -f python scripts/exp/ncu_shapes.py > gpurun_out/r2av_conv_ncu.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r2av_conv.ncu-rep --page raw --csv > gpurun_out/r2av_conv_ncu_full.csv 2>> gpurun_out/r2av_conv_ncu.log
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2av_conv_ncu_full.csv")))
keep = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "launch__registers_per_thread", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
idx = [i for i, h in enumerate(rows[0]) if h in keep or "tensor" in h]
csv.writer(open("gpurun_out/r2av_conv_ncu.csv", "w")).writerows([[r[i] for i in idx] for r in rows])
PY
rm -f gpurun_out/r2av_conv_ncu_full.csv
du -sh gpurun_out
