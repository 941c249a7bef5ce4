#!/bin/bash
# round 2, GPU call AP: second epilogue group in the generic implicit GEMM (Decoder / discriminator layers)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x > gpurun_out/r2ap_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2ap_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ap_bench.json 2> gpurun_out/r2ap_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ap_bench_again.json 2> gpurun_out/r2ap_bench_again.err; echo "bench again rc=$?"
for f in gpurun_out/r2ap_bench*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
fam=d["roofline"]["families"]
print(sys.argv[1], round(d["value"],1), "patches/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"], "| igemm128", round(fam["conv_igemm_tc<128>"]["ms_per_step"],2), round(fam["conv_igemm_tc<128>"]["tflops"]), "igemm64", round(fam["conv_igemm_tc<64>"]["ms_per_step"],2), "igemm32", round(fam["conv_igemm_tc<32>"]["ms_per_step"],2))
PY
done
