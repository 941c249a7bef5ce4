#!/bin/bash
# round 2, GPU call R: third lean epilogue group for 64->64 in the step, timeline gaps
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
SRCGAN_B200_SWEEP_GROUPS=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2r_bench_g2.json 2> gpurun_out/r2r_bench_g2.err; echo "bench g2 rc=$?"
timeout 300 python scripts/profile_step.py 64 > gpurun_out/r2r_profile_step.txt 2> gpurun_out/r2r_profile_step.err; echo "profile rc=$?"; sed -n 1,12p gpurun_out/r2r_profile_step.txt; sed -n '/^timeline/,$p' gpurun_out/r2r_profile_step.txt
for f in gpurun_out/r2r_bench*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"], "MHz")
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -3
