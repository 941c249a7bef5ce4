#!/bin/bash
# round 2, GPU call G: parity suite with planar concat buffers, A/B bench planar vs interleaved (10 steps each), step profile
set -u
mkdir -p gpurun_out
rm -f gpurun_out/reference_callers.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -8 gpurun_out/r2g_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
SRCGAN_B200_NO_PLANAR=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2g_bench_noplanar.json 2> gpurun_out/r2g_bench_noplanar.err; echo "bench noplanar rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2g_bench_again.json 2> gpurun_out/r2g_bench_again.err; echo "bench again rc=$?"
SRCGAN_B200_NO_PLANAR=1 SRCGAN_B200_NO_WGRAD_PAIRS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2g_bench_noplanar_nopairs.json 2> gpurun_out/r2g_bench_noplanar_nopairs.err; echo "bench noplanar nopairs rc=$?"
timeout 300 python scripts/profile_step.py 64 > gpurun_out/r2g_profile_step.txt 2> gpurun_out/r2g_profile_step.err; echo "profile rc=$?"; head -3 gpurun_out/r2g_profile_step.txt
du -sh gpurun_out
