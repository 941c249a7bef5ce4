#!/bin/bash
# round 2, GPU call K: fused pair - parity, in-kernel clock profile, variants
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "fused_pair or batchnorm" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2k_pytest.log
tail -15 gpurun_out/r2k_pytest.log
timeout 300 python scripts/exp/pair_bench.py 64 > gpurun_out/r2k_pair_bench.txt 2>&1; echo "pair_bench rc=$?"; cat gpurun_out/r2k_pair_bench.txt | grep -v "^{"
echo "---- profile (dbg 32)"
SRCGAN_B200_DBG=32 timeout 300 python scripts/exp/pair_bench.py 64 > gpurun_out/r2k_pair_prof.txt 2>&1; grep "pair mma (cta 0)\|pair epi (cta 0 warp 2\|pair epi (cta 0 warp 6\|pair epi (cta 0 warp 5" gpurun_out/r2k_pair_prof.txt | awk 'NR%23==1' | head -40
echo "---- no halo / no release (dbg 128)"
SRCGAN_B200_DBG=128 timeout 300 python scripts/exp/pair_bench.py 64 2>&1 | grep fused
echo "---- no stores (dbg 1)"
SRCGAN_B200_DBG=1 timeout 300 python scripts/exp/pair_bench.py 64 2>&1 | grep -v "^{"
