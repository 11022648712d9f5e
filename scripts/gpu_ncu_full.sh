#!/bin/bash
# ncu --set full captures of each hand-written kernel family (third, warm round of scripts/ncu_kernels.py)
# -> gpurun_out/ncu_<family>.ncu-rep; summarised into profiles/ with scripts/ncu_summary.py.
# usage: bash scripts/gpu_ncu_full.sh [family ...]   (default: all)
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() { fam=$1; regex=$2; skip=$3; cnt=$4; timeout 600 $NCU -k regex:$regex -s $skip -c $cnt -f -o gpurun_out/ncu_$fam python scripts/ncu_kernels.py $fam > gpurun_out/ncu_$fam.log 2>&1; }
want() { [ $# -eq 0 ] && return 0; for f in "$@"; do [ "$f" = "$FAM" ] && return 0; done; return 1; }
for FAM in gemm gn ln attn xattn fused; do
  want "$@" || continue
  case $FAM in
    gemm) cap gemm gemm_pair 24 12 ;;
    gn) cap gn "gn_(ring|slab)" 6 3 ;;
    ln) cap ln layernorm 6 3 ;;
    attn) cap attn temporal_attn 6 3 ;;
    xattn) cap xattn cross_attn 6 3 ;;
    fused) cap fused temporal_block 2 1 ;;
  esac
done
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_*.log
