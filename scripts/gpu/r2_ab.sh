#!/bin/bash
# round 2, GPU call AB: stacked wgrad with the phantom tap fed with zeros (power) - tests, per shape, step A/B
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "wgrad or conv_tc_full_size" 2>&1 | tail -2
echo "--- zeros"; timeout 300 python scripts/bench_wgrad.py tc 2>&1 | grep shape
echo "--- data"; SRCGAN_B200_WGRAD_PHANTOM=1 timeout 300 python scripts/bench_wgrad.py tc 2>&1 | grep shape
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ab_bench.json 2> gpurun_out/r2ab_bench.err; echo "bench rc=$?"
SRCGAN_B200_WGRAD_PHANTOM=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ab_bench_data.json 2> gpurun_out/r2ab_bench_data.err; echo "bench data rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ab_bench_again.json 2> gpurun_out/r2ab_bench_again.err; echo "bench again rc=$?"
SRCGAN_B200_WGRAD_PHANTOM=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ab_bench_data2.json 2> gpurun_out/r2ab_bench_data2.err; echo "bench data2 rc=$?"
for f in gpurun_out/r2ab_bench*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    fam=d["roofline"]["families"]["conv3x3_wgrad_stack_tc"]
    print(sys.argv[1], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"], "MHz", "wgrad", round(fam["ms_per_step"],1), "ms", round(fam["tflops"]), "TF", round(d["roofline"]["frac"],3))
except Exception as ex:
    print(sys.argv[1], "unreadable", ex)
PY
done
