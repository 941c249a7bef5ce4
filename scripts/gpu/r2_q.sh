#!/bin/bash
# round 2, GPU call Q: epilogue groups of the paired sweep
set -u
mkdir -p gpurun_out
for g in 2 3 4; do
echo "---- groups $g"
SRCGAN_B200_SWEEP_GROUPS=$g timeout 300 python scripts/exp/groups_bench.py 64 2>&1 | tee gpurun_out/r2q_groups_$g.txt
done
SRCGAN_B200_SWEEP_GROUPS=4 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv_tc" 2>&1 | tail -3
