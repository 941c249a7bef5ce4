#!/bin/bash
# round 2, GPU call B: full parity suite with buckets / batched repack, layout probe, per-layer incumbent table, bench
set -u
mkdir -p gpurun_out
rm -f gpurun_out/reference_callers.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -8 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
timeout 400 python scripts/exp/layout_probe.py > gpurun_out/r2b_layout_probe.json 2> gpurun_out/r2b_layout_probe.err; echo "probe rc=$?"
timeout 400 python scripts/bench_incumbent_layers.py > gpurun_out/r2b_incumbent_layers.json 2> gpurun_out/r2b_incumbent_layers.err; echo "layers rc=$?"
timeout 300 python scripts/prof_hbm.py > gpurun_out/r2b_hbm.json 2> gpurun_out/r2b_hbm.err; echo "hbm rc=$?"
du -sh gpurun_out
