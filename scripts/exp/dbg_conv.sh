#!/bin/bash
# component-elimination timing of the TC fprop kernels (SRCGAN_B200_DBG bits: 1 no stores, 2 no MMA, 4 no TMA loads,
# 8 no TMEM loads, 16 no shuffles)
for d in ${DBG_LIST:-0 1 2 4 3 5 6 7}; do
  echo "== DBG=$d"
  SRCGAN_B200_DBG=$d python scripts/bench_conv.py tc 2>&1 | grep shape
done
