#!/bin/bash
# round 2, GPU call AE (2 GPUs): where the 2-GPU step loses 10 ms - overlapped all-reduce vs flat, range mode on / off, dependent launch off
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { # name, env..., extra args after --
  name=$1; shift
  env "$@" timeout 600 $TR --master-port $((29700 + RANDOM % 200)) bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e $EXTRA > gpurun_out/r2ae_$name.json 2> gpurun_out/r2ae_$name.err; echo "$name rc=$?"
}
EXTRA="" run default A=1
EXTRA="--no-overlap" run flat A=1
EXTRA="" run noranges SRCGAN_B200_NO_SWEEP_RANGES=1
EXTRA="" run nopdl SRCGAN_B200_NO_PDL=1
for g in 0 1; do CUDA_VISIBLE_DEVICES=$g timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ae_1gpu_dev$g.json 2> gpurun_out/r2ae_1gpu_dev$g.err; echo "1gpu dev$g rc=$?"; done
for f in gpurun_out/r2ae_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["n_gpus"], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"])
except Exception as ex:
    print(sys.argv[1], "unreadable", ex)
PY
done
