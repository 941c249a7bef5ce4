#!/bin/bash
# round 2, GPU call AS: thin_in_tiled (FFMA kernel for <= 4 input channels) + pitch-8 thin tensors in the functional layer (cascade's
# 64 -> 3 prediction conv on tcgen05) - parity, A/B in the default and cascade benches; ncu of the 32 -> 32 remainder wgrad
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x > gpurun_out/r2as_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2as_pytest.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_kernels.py --deselect tests/test_gpu_models.py > gpurun_out/r2as_pytest_rest.log 2>&1; echo "pytest rest rc=$?"; tail -3 gpurun_out/r2as_pytest_rest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2as_bench.json 2> gpurun_out/r2as_bench.err; echo "bench rc=$?"
SRCGAN_B200_NO_THIN_TILED=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2as_bench_old.json 2> gpurun_out/r2as_bench_old.err; echo "bench old rc=$?"
timeout 600 python bench.py --workload cascade --by-shape --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2as_cascade_shapes.json 2> gpurun_out/r2as_cascade_shapes.err; echo "cascade rc=$?"
timeout 600 python bench.py --workload cascade --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2as_cascade.json 2> gpurun_out/r2as_cascade.err; echo "cascade rc=$?"
SRCGAN_B200_NO_THIN_TILED=1 SRCGAN_B200_NO_THIN_PITCH=1 timeout 600 python bench.py --workload cascade --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2as_cascade_old.json 2> gpurun_out/r2as_cascade_old.err; echo "cascade old rc=$?"
timeout 600 python bench.py --workload cascade_lab --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2as_cascade_lab.json 2> gpurun_out/r2as_cascade_lab.err; echo "cascade_lab rc=$?"
for f in gpurun_out/r2as_bench.json gpurun_out/r2as_bench_old.json gpurun_out/r2as_cascade.json gpurun_out/r2as_cascade_old.json gpurun_out/r2as_cascade_lab.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
fam=d["roofline"]["families"]
print(sys.argv[1], round(d["value"],1), d["unit"], round(d["ms_per_step"],2), "ms", d["clocks"]["sm_mhz"], "| simt fprop", round(fam.get("conv_fprop_simt",{}).get("ms_per_step",0),2), "dgrad", round(fam.get("conv_dgrad_simt",{}).get("ms_per_step",0),2), "wgrad", round(fam.get("conv_wgrad_simt",{}).get("ms_per_step",0),2))
PY
done
# what bounds the 32 -> 32 remainder wgrads: ncu --set full of both kernels
timeout 600 ncu --set full --clock-control none -k regex:'wgrad_stack|wgrad_r32' -c 8 -o /tmp/r2as_w32 -f python scripts/exp/wgrad32_bench.py > gpurun_out/r2as_ncu_w32.log 2>&1; echo "ncu rc=$?"
SRCGAN_B200_WGRAD_R32=1 timeout 600 ncu --set full --clock-control none -k regex:'wgrad_stack|wgrad_r32' -c 8 -o /tmp/r2as_w32r -f python scripts/exp/wgrad32_bench.py >> gpurun_out/r2as_ncu_w32.log 2>&1; echo "ncu rc=$?"
for t in r2as_w32 r2as_w32r; do ncu -i /tmp/$t.ncu-rep --page raw --csv > /tmp/$t.csv 2>/dev/null; python - /tmp/$t.csv gpurun_out/${t}_ncu.csv <<'PY'
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
keep=("ID","Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","dram__throughput.avg.pct_of_peak_sustained_elapsed","lts__t_bytes.sum","lts__t_sector_hit_rate.pct","lts__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__m_xbar2l1tex_read_bytes.sum","sm__throughput.avg.pct_of_peak_sustained_elapsed","sm__cycles_elapsed.max","sm__pipe_tensor_subpipe_cycles_active.avg.pct_of_peak_sustained_active","gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed","dram__cycles_active.avg.pct_of_peak_sustained_elapsed","fbpa__dram_read_throughput.avg.pct_of_peak_sustained_elapsed","lts__t_sectors_srcunit_tex_op_read.sum","lts__t_requests_srcunit_tex_op_read.sum", "dram__sectors_read.sum")
idx=[i for i,h in enumerate(rows[0]) if h in keep or "dram__" in h or "lts__t_sectors_op_read" in h or "fbp" in h]
csv.writer(open(sys.argv[2],"w")).writerows([[r[i] for i in idx] for r in rows])
PY
done
du -sh gpurun_out
