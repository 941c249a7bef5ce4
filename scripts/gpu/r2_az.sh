#!/bin/bash
# round 2, GPU call AZ: split-K reduce with eight loads in flight per thread (same summation order) - wgrad tests, A/B
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "wgrad" > gpurun_out/r2az_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2az_pytest.log
SRCGAN_B200_NO_PDL=1 timeout 300 python scripts/profile_step.py 64 2>/dev/null | grep -E "wgrad_reduce|batch 64" | head -3
SRCGAN_B200_NO_PDL=1 SRCGAN_B200_REDUCE_ILP4=1 timeout 300 python scripts/profile_step.py 64 2>/dev/null | grep -E "wgrad_reduce|batch 64" | head -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2az_bench.json 2> gpurun_out/r2az_bench.err; echo "bench rc=$?"
SRCGAN_B200_REDUCE_ILP4=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2az_bench_old.json 2> gpurun_out/r2az_bench_old.err; echo "bench old rc=$?"
for f in gpurun_out/r2az_bench.json gpurun_out/r2az_bench_old.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
fam=d["roofline"]["families"]
print(sys.argv[1], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"], "| wgrad_stack", round(fam["conv3x3_wgrad_stack_tc"]["ms_per_step"],2))
PY
done
