#!/bin/bash
set -u
mkdir -p gpurun_out
for i in 1 2; do
echo "--- zeros"; timeout 120 python scripts/exp/wgrad32_bench.py 64
echo "--- data"; SRCGAN_B200_WGRAD_PHANTOM=1 timeout 120 python scripts/exp/wgrad32_bench.py 64
done
echo "--- zeros"; timeout 300 python scripts/bench_wgrad.py tc 2>&1 | grep shape
echo "--- data"; SRCGAN_B200_WGRAD_PHANTOM=1 timeout 300 python scripts/bench_wgrad.py tc 2>&1 | grep shape
