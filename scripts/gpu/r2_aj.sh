#!/bin/bash
# round 2, GPU call AJ: final build - full GPU suite, smoke, default bench
set -u
mkdir -p gpurun_out
rm -f gpurun_out/reference_callers.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2aj_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2aj_pytest.log
tail -3 gpurun_out/r2aj_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2aj_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2aj_smoke.log
timeout 900 python bench.py > gpurun_out/r2aj_bench.json 2> gpurun_out/r2aj_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2aj_bench.json").read().strip().splitlines()[-1])
e=d.get("e2e") or {}
print(round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", "e2e", round(e.get("value",0),1), d["clocks"], d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["gpu_launches"], d.get("cpu_baseline"))
PY
