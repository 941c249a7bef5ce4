#!/bin/bash
# round 2, GPU call E (2 GPUs): data-parallel correctness + overlapped vs flat all-reduce bench
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/dp_check.py > gpurun_out/r2e_dp_check.json 2> gpurun_out/r2e_dp_check.err; echo "dp_check rc=$?"; tail -3 gpurun_out/r2e_dp_check.json; tail -5 gpurun_out/r2e_dp_check.err
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2e_bench_2gpu.json 2> gpurun_out/r2e_bench_2gpu.err; echo "bench2 rc=$?"
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-overlap --no-e2e > gpurun_out/r2e_bench_2gpu_flat.json 2> gpurun_out/r2e_bench_2gpu_flat.err; echo "bench2 flat rc=$?"
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_1gpu.json 2> gpurun_out/r2e_bench_1gpu.err; echo "bench1 rc=$?"
tail -c 600 gpurun_out/r2e_bench_2gpu.json
