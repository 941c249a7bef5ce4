#!/bin/bash
# round 2, GPU call P: fused pair in the step - full GPU suite, A/B bench (fused pairs on / off / two issuers), step profile
set -u
mkdir -p gpurun_out
rm -f gpurun_out/reference_callers.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -8 gpurun_out/r2p_pytest.log
timeout 300 python scripts/exp/pair_bench.py 64 2>&1 | grep -v "^{" | tee gpurun_out/r2p_pair_bench.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"
SRCGAN_B200_NO_PAIR_FPROP=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2p_bench_nofuse.json 2> gpurun_out/r2p_bench_nofuse.err; echo "bench nofuse rc=$?"
SRCGAN_B200_PAIR_ISSUERS=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2p_bench_2iss.json 2> gpurun_out/r2p_bench_2iss.err; echo "bench 2iss rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2p_bench_again.json 2> gpurun_out/r2p_bench_again.err; echo "bench again rc=$?"
timeout 300 python scripts/profile_step.py 64 > gpurun_out/r2p_profile_step.txt 2> gpurun_out/r2p_profile_step.err; echo "profile rc=$?"; sed -n 1,30p gpurun_out/r2p_profile_step.txt
for f in gpurun_out/r2p_bench*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"], "MHz")
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
