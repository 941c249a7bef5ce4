#!/bin/bash
# round 2, GPU call AC: the other workloads (BASELINE configs 3-5) on the final build
set -u
mkdir -p gpurun_out
for wl in cascade cascade_lab eval; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/r2ac_bench_$wl.json 2> gpurun_out/r2ac_bench_$wl.err; echo "bench $wl rc=$?"
  python - gpurun_out/r2ac_bench_$wl.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d.get("e2e") or {}
    print(sys.argv[1], round(d["value"],1), d["unit"], round(d["ms_per_step"],1), "ms", "e2e", round(e.get("value",0),1), d.get("cpu_baseline",{}).get("value"))
except Exception as ex:
    print(sys.argv[1], "unreadable", ex)
PY
done
