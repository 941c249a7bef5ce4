#!/bin/bash
# round 2, GPU call W: range-mode scheduling of the paired sweep (all 74 CTA pairs busy)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv_tc" > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2w_pytest.log
tail -4 gpurun_out/r2w_pytest.log
echo "--- ranges"; timeout 300 python scripts/exp/groups_bench.py 64 2>&1 | tee gpurun_out/r2w_ranges.txt
echo "--- units"; SRCGAN_B200_NO_SWEEP_RANGES=1 timeout 300 python scripts/exp/groups_bench.py 64 2>&1 | tee gpurun_out/r2w_units.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"
SRCGAN_B200_NO_SWEEP_RANGES=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2w_bench_units.json 2> gpurun_out/r2w_bench_units.err; echo "bench units rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2w_bench_again.json 2> gpurun_out/r2w_bench_again.err; echo "bench again rc=$?"
for f in gpurun_out/r2w_bench*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"], "MHz", d["roofline"]["kernel"], round(d["roofline"]["frac"],3))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
