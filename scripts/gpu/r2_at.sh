#!/bin/bash
# round 2, GPU call AT: conv3x3_wgrad_r32_tc with dY through TMA (four taps in M, 8 MMAs per tile) - parity, per-launch time, bench A/B
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "wgrad" > gpurun_out/r2at_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2at_pytest.log
rm -f gpurun_out/r2at_wgrad32.txt
for m in tma lsu 0; do echo "mode $m" | tee -a gpurun_out/r2at_wgrad32.txt; SRCGAN_B200_WGRAD_R32=$m timeout 200 python scripts/exp/wgrad32_bench.py 2>&1 | tee -a gpurun_out/r2at_wgrad32.txt; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2at_bench.json 2> gpurun_out/r2at_bench.err; echo "bench rc=$?"
SRCGAN_B200_WGRAD_R32=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2at_bench_old.json 2> gpurun_out/r2at_bench_old.err; echo "bench old rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2at_bench_again.json 2> gpurun_out/r2at_bench_again.err; echo "bench rc=$?"
for f in gpurun_out/r2at_bench.json gpurun_out/r2at_bench_old.json gpurun_out/r2at_bench_again.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
fam=d["roofline"]["families"]
print(sys.argv[1], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"], "| wgrad_stack", round(fam["conv3x3_wgrad_stack_tc"]["ms_per_step"],2))
PY
done
