#!/bin/bash
# round 2, GPU call L: fused pair - parity after the ragged-lane fix, latency trace, variants
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "fused_pair" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2l_pytest.log
tail -8 gpurun_out/r2l_pytest.log
for d in 32 160 416 928; do
echo "---- profile (dbg $d)"
SRCGAN_B200_DBG=$d timeout 300 python scripts/exp/pair_bench.py 64 > gpurun_out/r2l_pair_prof_$d.txt 2>&1
grep "fused\|separate" gpurun_out/r2l_pair_prof_$d.txt | grep -v "^{"
grep "pair trace" gpurun_out/r2l_pair_prof_$d.txt | sed -n '97,104p'
grep "pair mma (cta 0)" gpurun_out/r2l_pair_prof_$d.txt | sed -n '7p;30p'
grep "pair epi (cta 0 warp 4" gpurun_out/r2l_pair_prof_$d.txt | sed -n '7p;30p'
done
