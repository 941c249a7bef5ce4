#!/bin/bash
# round 2, GPU call Z: per-phase clocks of the fused pair's epilogue groups (instrumented build, not committed)
set -u
mkdir -p gpurun_out
SRCGAN_B200_DBG=32 timeout 300 python scripts/exp/pair_bench.py 64 > gpurun_out/r2z_pair_prof.txt 2>&1
grep -c "pair epi" gpurun_out/r2z_pair_prof.txt
# one launch = 32 epi lines (2 ctas x 16 warps); take a block in the middle of the first fused config (64+96) and of the 128+160 one
grep "pair epi\|pair mma" gpurun_out/r2z_pair_prof.txt | sed -n '341,378p'
echo ----
grep "pair epi\|pair mma" gpurun_out/r2z_pair_prof.txt | sed -n '2041,2078p'
