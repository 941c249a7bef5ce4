#!/bin/bash
# round 2, GPU call AD (2 GPUs): data-parallel check and bench on the final kernels (fused pairs, dependent launches)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29611 scripts/dp_check.py > gpurun_out/r2ad_dp_check.json 2> gpurun_out/r2ad_dp_check.err; echo "dp_check rc=$?"; tail -3 gpurun_out/r2ad_dp_check.json
timeout 600 $TR --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2ad_bench_2gpu.json 2> gpurun_out/r2ad_bench_2gpu.err; echo "bench2 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2ad_bench_1gpu.json 2> gpurun_out/r2ad_bench_1gpu.err; echo "bench1 rc=$?"
for f in gpurun_out/r2ad_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d.get("e2e") or {}
    print(sys.argv[1], d["n_gpus"], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", "e2e", round(e.get("value",0),1), d["config"].get("allreduce"))
except Exception as ex:
    print(sys.argv[1], "unreadable", ex)
PY
done
