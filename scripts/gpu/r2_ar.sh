#!/bin/bash
# round 2, GPU call AR: conv3x3_wgrad_r32_tc (32 -> 32 remainder wgrads: cp.async dY producer, four taps in M) - parity, per-launch
# time against the stacked <32> kernel, bench A/B; per-shape profile of the cascade workload
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "wgrad" > gpurun_out/r2ar_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ar_pytest.log
timeout 200 python scripts/exp/wgrad32_bench.py > gpurun_out/r2ar_wgrad32.txt 2>&1; echo "rc=$?"; cat gpurun_out/r2ar_wgrad32.txt
SRCGAN_B200_NO_WGRAD_R32=1 timeout 200 python scripts/exp/wgrad32_bench.py > gpurun_out/r2ar_wgrad32_old.txt 2>&1; cat gpurun_out/r2ar_wgrad32_old.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ar_bench.json 2> gpurun_out/r2ar_bench.err; echo "bench rc=$?"
SRCGAN_B200_NO_WGRAD_R32=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ar_bench_old.json 2> gpurun_out/r2ar_bench_old.err; echo "bench old rc=$?"
timeout 600 python bench.py --workload cascade --by-shape --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ar_cascade_shapes.json 2> gpurun_out/r2ar_cascade_shapes.err; echo "cascade rc=$?"
for f in gpurun_out/r2ar_bench.json gpurun_out/r2ar_bench_old.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
fam=d["roofline"]["families"]
print(sys.argv[1], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"], "| wgrad_stack", round(fam["conv3x3_wgrad_stack_tc"]["ms_per_step"],2))
PY
done
