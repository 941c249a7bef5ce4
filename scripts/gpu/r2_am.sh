#!/bin/bash
# round 2, GPU call AM (8 GPUs): weak-scaling bench on the final build
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29911 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2am_bench_8gpu.json 2> gpurun_out/r2am_bench_8gpu.err; echo "bench8 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2am_bench_8gpu.json").read().strip().splitlines()[-1])
    e=d.get("e2e") or {}
    print(d["n_gpus"], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", "e2e", round(e.get("value",0),1), d["clocks"], d["config"].get("allreduce"))
except Exception as ex:
    print("unreadable", ex)
PY
tail -3 gpurun_out/r2am_bench_8gpu.err
