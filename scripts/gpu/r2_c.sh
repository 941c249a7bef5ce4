#!/bin/bash
# round 2, GPU call C: parity suite (tall-image mode, noise-floor full-size test), bench of all four workloads
set -u
mkdir -p gpurun_out
rm -f gpurun_out/reference_callers.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -8 gpurun_out/r2c_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
SRCGAN_B200_NO_TALL=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c_bench_notall.json 2> gpurun_out/r2c_bench_notall.err; echo "bench notall rc=$?"
for wl in cascade cascade_lab eval; do
  timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 > gpurun_out/r2c_bench_$wl.json 2> gpurun_out/r2c_bench_$wl.err; echo "bench $wl rc=$?"
done
timeout 300 python scripts/prof_hbm.py > gpurun_out/r2c_hbm.json 2> gpurun_out/r2c_hbm.err; echo "hbm rc=$?"
du -sh gpurun_out
