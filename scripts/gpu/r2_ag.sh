#!/bin/bash
# round 2, GPU call AG: 64-byte L2 promotion for the 32-channel slices of the wgrad tensor maps
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "wgrad" 2>&1 | tail -2
echo "--- 64B"; timeout 120 python scripts/exp/wgrad32_bench.py 64
echo "--- 128B"; SRCGAN_B200_NO_L2_64B=1 timeout 120 python scripts/exp/wgrad32_bench.py 64
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ag_bench.json 2> gpurun_out/r2ag_bench.err; echo "bench rc=$?"
SRCGAN_B200_NO_L2_64B=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ag_bench_128.json 2> gpurun_out/r2ag_bench_128.err; echo "bench 128 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ag_bench_again.json 2> gpurun_out/r2ag_bench_again.err; echo "bench again rc=$?"
SRCGAN_B200_NO_L2_64B=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ag_bench_128b.json 2> gpurun_out/r2ag_bench_128b.err; echo "bench 128b rc=$?"
for f in gpurun_out/r2ag_bench*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    fam=d["roofline"]["families"]["conv3x3_wgrad_stack_tc"]
    print(sys.argv[1], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"], "MHz", "wgrad", round(fam["ms_per_step"],1), "ms", round(d["roofline"]["frac"],3))
except Exception as ex:
    print(sys.argv[1], "unreadable", ex)
PY
done
