#!/bin/bash
# round 2, GPU call M: fused pair v2 (4 epilogue groups, lag 3, st.async halo)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "fused_pair" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2m_pytest.log
tail -8 gpurun_out/r2m_pytest.log
timeout 300 python scripts/exp/pair_bench.py 64 > gpurun_out/r2m_pair_bench.txt 2>&1; echo "pair_bench rc=$?"; grep -v "^{" gpurun_out/r2m_pair_bench.txt
for d in 32 128; do
echo "---- (dbg $d)"
SRCGAN_B200_DBG=$d timeout 300 python scripts/exp/pair_bench.py 64 > gpurun_out/r2m_pair_prof_$d.txt 2>&1
grep "fused" gpurun_out/r2m_pair_prof_$d.txt | grep -v "^{"
grep "pair mma (cta 0)" gpurun_out/r2m_pair_prof_$d.txt | sed -n '7p;30p;53p;76p'
done
