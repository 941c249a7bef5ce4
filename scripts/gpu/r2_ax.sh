#!/bin/bash
# round 2, GPU call AX: 3x3 weight gradient of thin-output layers (64 -> 3) on the stacked kernel - parity, bench A/B
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "thin or wgrad or fprop_dgrad" > gpurun_out/r2ax_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ax_pytest.log
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2ax_pytest_models.log 2>&1; echo "pytest models rc=$?"; tail -3 gpurun_out/r2ax_pytest_models.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ax_bench.json 2> gpurun_out/r2ax_bench.err; echo "bench rc=$?"
SRCGAN_B200_NO_WSTACK_THIN=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ax_bench_old.json 2> gpurun_out/r2ax_bench_old.err; echo "bench old rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ax_bench_again.json 2> gpurun_out/r2ax_bench_again.err; echo "bench rc=$?"
for f in gpurun_out/r2ax_bench.json gpurun_out/r2ax_bench_old.json gpurun_out/r2ax_bench_again.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
fam=d["roofline"]["families"]
print(sys.argv[1], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"], "| wgrad_stack", round(fam["conv3x3_wgrad_stack_tc"]["ms_per_step"],2), "wgrad_tc", round(fam["conv_wgrad_tc"]["ms_per_step"],2), "halo64", round(fam.get("conv3x3_wgrad_halo_tc<64>",{}).get("ms_per_step",0),2))
PY
done
