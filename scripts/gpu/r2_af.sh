#!/bin/bash
# round 2, GPU call AF (2 GPUs): bucket all-reduce launched in the step pre-hook (new default) vs from inside backward
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29811 scripts/dp_check.py > gpurun_out/r2af_dp_check.json 2> gpurun_out/r2af_dp_check.err; echo "dp_check rc=$?"; tail -1 gpurun_out/r2af_dp_check.json
DP_CHECK_OVERLAP=1 timeout 600 $TR --master-port 29812 scripts/dp_check.py > gpurun_out/r2af_dp_check_overlap.json 2> gpurun_out/r2af_dp_check_overlap.err; echo "dp_check overlap rc=$?"; tail -1 gpurun_out/r2af_dp_check_overlap.json
timeout 600 $TR --master-port 29813 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2af_bench_2gpu.json 2> gpurun_out/r2af_bench_2gpu.err; echo "bench2 rc=$?"
timeout 600 $TR --master-port 29814 bench.py --gpus 2 --steps 10 --warmup 3 --overlap --no-e2e > gpurun_out/r2af_bench_2gpu_overlap.json 2> gpurun_out/r2af_bench_2gpu_overlap.err; echo "bench2 overlap rc=$?"
for g in 0 1; do CUDA_VISIBLE_DEVICES=$g timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2af_bench_1gpu_dev$g.json 2> gpurun_out/r2af_bench_1gpu_dev$g.err; echo "1gpu dev$g rc=$?"; done
for f in gpurun_out/r2af_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d.get("e2e") or {}
    print(sys.argv[1], d["n_gpus"], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", "e2e", round(e.get("value",0),1), (d["config"].get("allreduce") or {}).get("pre_hook_launches"), (d["config"].get("allreduce") or {}).get("overlapped_launches"))
except Exception as ex:
    print(sys.argv[1], "unreadable", ex)
PY
done
