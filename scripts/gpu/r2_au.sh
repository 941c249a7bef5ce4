#!/bin/bash
# round 2, GPU call AU: phantom-free third tap of the stacked wgrad kernel (T22: 96 + 64 clk per k-step instead of 96 + 96) - parity,
# per-shape times, bench A/B
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "wgrad" > gpurun_out/r2au_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2au_pytest.log
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2au_pytest_models.log 2>&1; echo "pytest models rc=$?"; tail -3 gpurun_out/r2au_pytest_models.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2au_bench.json 2> gpurun_out/r2au_bench.err; echo "bench rc=$?"
SRCGAN_B200_WGRAD_T22=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2au_bench_old.json 2> gpurun_out/r2au_bench_old.err; echo "bench old rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2au_bench_again.json 2> gpurun_out/r2au_bench_again.err; echo "bench rc=$?"
for f in gpurun_out/r2au_bench.json gpurun_out/r2au_bench_old.json gpurun_out/r2au_bench_again.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
fam=d["roofline"]["families"]
w=fam["conv3x3_wgrad_stack_tc"]
print(sys.argv[1], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"], "| wgrad_stack", round(w["ms_per_step"],2), "ms", round(w["tflops"]), "TFLOP/s")
PY
done
timeout 300 python scripts/profile_shapes.py > gpurun_out/r2au_profile_shapes.txt 2> gpurun_out/r2au_profile_shapes.err; echo "shapes rc=$?"; grep wgrad gpurun_out/r2au_profile_shapes.txt | head -8
