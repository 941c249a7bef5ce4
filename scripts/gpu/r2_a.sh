#!/bin/bash
# round 2, GPU call A: parity suite, bench (ours / incumbent), per-layer incumbent table, HBM kernel timing + ncu
set -u
mkdir -p gpurun_out
rm -f gpurun_out/reference_callers.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl torch-gpu --batch 16 --steps 3 --warmup 2 > gpurun_out/r2a_incumbent.json 2> gpurun_out/r2a_incumbent.err; echo "incumbent rc=$?"
timeout 300 python scripts/bench_incumbent_layers.py > gpurun_out/r2a_incumbent_layers.json 2> gpurun_out/r2a_incumbent_layers.err; echo "layers rc=$?"
timeout 300 python scripts/prof_hbm.py > gpurun_out/r2a_hbm.json 2> gpurun_out/r2a_hbm.err; echo "hbm rc=$?"
# ncu: the report stays on the box (it can exceed the 64 MiB return limit); only the per-kernel CSV comes back
timeout 900 ncu --set full --clock-control none -c 120 \
  -k regex:'bn_|loss_|ssim|ae_k|minmax|rgb2lab|lab2rgb|upsample|nchw|nhwc|colsum|reduce|add_|eval_metrics' \
  -o /tmp/r2a_hbm -f python scripts/prof_hbm.py --once --n 16 > gpurun_out/r2a_hbm_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r2a_hbm.ncu-rep --page raw --csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active \
  > gpurun_out/r2a_hbm_ncu.csv 2>> gpurun_out/r2a_hbm_ncu.log
du -sh gpurun_out; ls -la gpurun_out | tail -14
