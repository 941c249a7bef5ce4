#!/bin/bash
# round 2, GPU call O: fused pair v3 (second MMA issuer for the x_k part)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "fused_pair" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2o_pytest.log
tail -4 gpurun_out/r2o_pytest.log
timeout 300 python scripts/exp/pair_bench.py 64 2>&1 | grep -v "^{"
echo "---- dbg 32 profile"
SRCGAN_B200_DBG=32 timeout 300 python scripts/exp/pair_bench.py 64 > gpurun_out/r2o_pair_prof.txt 2>&1
grep "pair mma (cta 0)" gpurun_out/r2o_pair_prof.txt | sed -n '7p;30p;53p;76p'
