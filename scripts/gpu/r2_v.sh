#!/bin/bash
# round 2, GPU call V: ring-split issuers in the step (A/B on one box), full GPU suite
set -u
mkdir -p gpurun_out
rm -f gpurun_out/reference_callers.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
SRCGAN_B200_PAIR_ISSUERS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2v_bench_iss1.json 2> gpurun_out/r2v_bench_iss1.err; echo "bench iss1 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2v_bench_again.json 2> gpurun_out/r2v_bench_again.err; echo "bench again rc=$?"
for f in gpurun_out/r2v_bench*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"], "MHz", d["roofline"]["kernel"], round(d["roofline"]["frac"],3))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
tail -4 gpurun_out/r2v_pytest.log
