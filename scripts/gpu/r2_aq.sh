#!/bin/bash
# round 2, GPU call AQ (final build): whole GPU suite + smoke, per-shape profile of one step, ncu launch list of the bench command
# (gpu__time_duration only), default bench line
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2aq_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2aq_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2aq_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2aq_smoke.log
timeout 300 python scripts/profile_shapes.py > gpurun_out/r2aq_profile_shapes.txt 2> gpurun_out/r2aq_profile_shapes.err; echo "shapes rc=$?"
timeout 600 python bench.py > gpurun_out/r2aq_bench.json 2> gpurun_out/r2aq_bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2aq_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2aq_ncu_bench.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r2aq_launches.csv
gzip -f gpurun_out/r2aq_launches.csv
du -sh gpurun_out
