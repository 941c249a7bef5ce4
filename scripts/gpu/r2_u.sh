#!/bin/bash
# round 2, GPU call U: fused pair - one MMA-issuing warp per ring
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "fused_pair" > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2u_pytest.log
tail -4 gpurun_out/r2u_pytest.log
echo "--- issuers 2 (ring split)"; timeout 300 python scripts/exp/pair_bench.py 64 2>&1 | grep -v "^{" | tee gpurun_out/r2u_pair_bench.txt
echo "--- issuers 1"; SRCGAN_B200_PAIR_ISSUERS=1 timeout 300 python scripts/exp/pair_bench.py 64 2>&1 | grep "fused"
echo "---- dbg 32 profile"
SRCGAN_B200_DBG=32 timeout 300 python scripts/exp/pair_bench.py 64 > gpurun_out/r2u_pair_prof.txt 2>&1
grep "pair mma warp . (cta 0)" gpurun_out/r2u_pair_prof.txt | sed -n '13,14p;59,60p;105,106p;151,152p'
