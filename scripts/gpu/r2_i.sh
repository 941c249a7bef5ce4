#!/bin/bash
# round 2, GPU call I: host-side profile of the step, BN mask recompute test
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "batchnorm or wgrad or planar" > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -4 gpurun_out/r2i_pytest.log
PROFILE_STEP_HOST=1 timeout 300 python scripts/profile_step.py 64 > gpurun_out/r2i_profile_step.txt 2> gpurun_out/r2i_profile_step.err; echo "profile rc=$?"; head -60 gpurun_out/r2i_profile_step.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
