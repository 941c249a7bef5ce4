#!/bin/bash
# round 2, GPU call Y: programmatic dependent launch on the sweep / pair / stacked wgrad / reduce kernels
set -u
mkdir -p gpurun_out
rm -f gpurun_out/reference_callers.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log
tail -4 gpurun_out/r2y_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"
SRCGAN_B200_NO_PDL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2y_bench_nopdl.json 2> gpurun_out/r2y_bench_nopdl.err; echo "bench nopdl rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2y_bench_again.json 2> gpurun_out/r2y_bench_again.err; echo "bench again rc=$?"
for f in gpurun_out/r2y_bench*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"], "MHz", d["roofline"]["kernel"], round(d["roofline"]["frac"],3))
except Exception as ex:
    print(sys.argv[1], "unreadable", ex)
PY
done
timeout 300 python scripts/profile_step.py 64 > gpurun_out/r2y_profile_step.txt 2> gpurun_out/r2y_profile_step.err; echo "profile rc=$?"; sed -n 1,4p gpurun_out/r2y_profile_step.txt; sed -n '/^timeline/,/^largest/p' gpurun_out/r2y_profile_step.txt
