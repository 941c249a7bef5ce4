#!/bin/bash
# round 2, GPU call H: parity suite, latency-built split-K reduce, A/B of the wgrad pairing, step profile
set -u
mkdir -p gpurun_out
rm -f gpurun_out/reference_callers.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -8 gpurun_out/r2h_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
SRCGAN_B200_NO_WGRAD_PAIRS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2h_bench_nopairs.json 2> gpurun_out/r2h_bench_nopairs.err; echo "bench nopairs rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2h_bench_again.json 2> gpurun_out/r2h_bench_again.err; echo "bench again rc=$?"
timeout 300 python scripts/profile_step.py 64 > gpurun_out/r2h_profile_step.txt 2> gpurun_out/r2h_profile_step.err; echo "profile rc=$?"; head -3 gpurun_out/r2h_profile_step.txt
timeout 600 python bench.py --workload cascade --steps 3 --warmup 3 > gpurun_out/r2h_bench_cascade.json 2> gpurun_out/r2h_bench_cascade.err; echo "bench cascade rc=$?"
du -sh gpurun_out
