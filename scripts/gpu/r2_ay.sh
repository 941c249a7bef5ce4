#!/bin/bash
# round 2, GPU call AY: whole GPU suite, smoke and the default bench line on the very last build
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2ay_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2ay_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ay_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2ay_smoke.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/r2ay_bench.json 2> gpurun_out/r2ay_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2ay_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print(round(d["value"],1), d["unit"], round(d["ms_per_step"],2), "ms | e2e", round(d["e2e"]["value"],1), "| clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "| roofline", r["kernel"], round(r["frac"],3), round(r["traffic"]/1e6), "MB | cpu", round(d["cpu_baseline"]["value"],3), d["cpu_baseline"]["kind"], "| launches", d["gpu_launches"])
PY
