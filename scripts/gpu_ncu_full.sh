#!/bin/bash
# ncu --set full captures of each hand-written kernel family (third, warm round of scripts/ncu_kernels.py)
# -> gpurun_out/ncu_<family>.ncu-rep; read them here with scripts/ncu_summary.py / `ncu -i ... --page source --print-source sass`.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() { fam=$1; regex=$2; skip=$3; cnt=$4; timeout 600 $NCU -k regex:$regex -s $skip -c $cnt -f -o gpurun_out/ncu_$fam python scripts/ncu_kernels.py $fam > gpurun_out/ncu_$fam.log 2>&1; }
cap gemm gemm_pair 16 8
cap gn "gn_(ring|slab)" 4 2
cap ln layernorm 4 2
cap attn temporal_attn 4 2
cap xattn cross_attn 4 2
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_*.log
