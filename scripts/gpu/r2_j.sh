#!/bin/bash
# round 2, GPU call J: fused dense-block layer pair - parity tests, micro-benchmark, step bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fused_pair or batchnorm" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2j_pytest.log
tail -30 gpurun_out/r2j_pytest.log
timeout 300 python scripts/exp/pair_bench.py 64 > gpurun_out/r2j_pair_bench.txt 2> gpurun_out/r2j_pair_bench.err; echo "pair_bench rc=$?"; cat gpurun_out/r2j_pair_bench.txt; tail -5 gpurun_out/r2j_pair_bench.err
