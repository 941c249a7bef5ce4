#!/bin/bash
set -u
mkdir -p gpurun_out
PROFILE_STEP_HOST=1 SRCGAN_B200_NO_PDL=1 timeout 300 python scripts/profile_step.py 64 > gpurun_out/r2ak_profile_step.txt 2> gpurun_out/r2ak_profile_step.err; echo "profile rc=$?"; sed -n 1,45p gpurun_out/r2ak_profile_step.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2ak_bench.json 2> gpurun_out/r2ak_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2ak_bench.json").read().strip().splitlines()[-1])
e=d.get("e2e") or {}
print(round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", "e2e", round(e.get("value",0),1))
PY
