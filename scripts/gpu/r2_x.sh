#!/bin/bash
# round 2, GPU call X: full default bench (e2e through the prefetching feed, cpu baseline), WSTACK32 A/B, smoke, reference arm
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2x_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2x_smoke.log
timeout 300 python -m pytest tests/test_gpu_models.py -m gpu -q -k "prefetcher" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"
SRCGAN_B200_WSTACK32=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2x_bench_wstack32.json 2> gpurun_out/r2x_bench_wstack32.err; echo "bench wstack32 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2x_bench_again.json 2> gpurun_out/r2x_bench_again.err; echo "bench again rc=$?"
for f in gpurun_out/r2x_bench*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d.get("e2e") or {}
    print(sys.argv[1], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", "e2e", round(e.get("value",0),1), d["clocks"]["sm_mhz"], "MHz", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d.get("cpu_baseline"))
except Exception as ex:
    print(sys.argv[1], "unreadable", ex)
PY
done
timeout 300 python scripts/profile_step.py 64 > gpurun_out/r2x_profile_step.txt 2> gpurun_out/r2x_profile_step.err; echo "profile rc=$?"; sed -n 1,40p gpurun_out/r2x_profile_step.txt
