#!/bin/bash
# round 2, GPU call AW (2 GPUs): the operator layer's bf16 thin-tensor test, data-parallel check and 2-GPU bench on the last build
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -x -k "thin_tensors or zoo or cascade" > gpurun_out/r2aw_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2aw_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29811 scripts/dp_check.py > gpurun_out/r2aw_dp_check.json 2> gpurun_out/r2aw_dp_check.err; echo "dp_check rc=$?"; tail -1 gpurun_out/r2aw_dp_check.json | cut -c1-600
timeout 600 $TR --master-port 29813 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2aw_bench_2gpu.json 2> gpurun_out/r2aw_bench_2gpu.err; echo "bench2 rc=$?"
for g in 0 1; do CUDA_VISIBLE_DEVICES=$g timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2aw_bench_1gpu_dev$g.json 2> gpurun_out/r2aw_bench_1gpu_dev$g.err; echo "1gpu dev$g rc=$?"; done
for f in gpurun_out/r2aw_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d.get("e2e") or {}
    print(sys.argv[1], d["n_gpus"], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", "e2e", round(e.get("value",0),1), d["clocks"]["sm_mhz"])
except Exception as ex:
    print(sys.argv[1], "unreadable", ex)
PY
done
