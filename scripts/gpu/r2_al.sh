#!/bin/bash
# round 2, GPU call AL: fused pair - lag of the x_k stream (2 / 3 / 4) and XK ring depth (4 / 8) - experiment builds
set -u
mkdir -p gpurun_out
L=srcgan_b200/lib
cp $L/libsrcgan_b200.so /tmp/lib_default.so
for v in default lag2 lag4 nxk8 lag4nxk8 default; do
  if [ $v = default ]; then cp /tmp/lib_default.so $L/libsrcgan_b200.so; else cp $L/libsrcgan_b200_$v.so $L/libsrcgan_b200.so; fi
  echo "---- $v"
  timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "fused_pair" 2>&1 | tail -1
  timeout 300 python scripts/exp/pair_bench.py 64 2>&1 | grep "fused" | grep -v "^{"
done
cp /tmp/lib_default.so $L/libsrcgan_b200.so
