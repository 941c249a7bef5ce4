#!/bin/bash
# round 2, GPU call N: fused pair v2 + barrier peeks, L2 prefetch distance sweep, plain xk wait
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "fused_pair" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2n_pytest.log
tail -4 gpurun_out/r2n_pytest.log
for pfd in 0 4 8 16; do
echo "---- pfd $pfd"
SRCGAN_B200_PAIR_PFD=$pfd timeout 300 python scripts/exp/pair_bench.py 64 2>&1 | grep "fused"
done
echo "---- dbg 256 (plain xk wait), pfd 8"
SRCGAN_B200_DBG=256 SRCGAN_B200_PAIR_PFD=8 timeout 300 python scripts/exp/pair_bench.py 64 2>&1 | grep "fused"
echo "---- dbg 32 profile, pfd 8"
SRCGAN_B200_DBG=32 SRCGAN_B200_PAIR_PFD=8 timeout 300 python scripts/exp/pair_bench.py 64 > gpurun_out/r2n_pair_prof.txt 2>&1
grep "pair mma (cta 0)" gpurun_out/r2n_pair_prof.txt | sed -n '7p;30p;53p;76p'
