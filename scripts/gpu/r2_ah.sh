#!/bin/bash
set -u
mkdir -p gpurun_out
SRCGAN_B200_NO_PDL=1 timeout 300 python scripts/profile_step.py 64 > gpurun_out/r2ah_profile_step.txt 2> gpurun_out/r2ah_profile_step.err; echo "profile rc=$?"; sed -n 1,60p gpurun_out/r2ah_profile_step.txt
