#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fmha_kernel -s 2 -c 1 -f -o gpurun_out/ncu_fmha python scripts/fmha_lab.py > gpurun_out/ncu_fmha.log 2>&1
tail -3 gpurun_out/ncu_fmha.log; ls -la gpurun_out/ncu_fmha.ncu-rep
