#!/bin/bash
# round 2, GPU call S: reduce + bias finalisation in one launch (tests), idle-SM experiment (64 vs 74 images per launch)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "wgrad or conv or bias" 2>&1 | tail -3
for s in "192 64" "64 64" "64 32"; do timeout 120 python scripts/exp/conv_scale.py $s 64 74 2>&1 | tail -2; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2s_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"], "MHz", d["gpu_launches"])
PY
