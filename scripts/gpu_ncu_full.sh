#!/bin/bash
# ncu --set full captures of each kernel family (third, warm round of scripts/ncu_kernels.py) -> gpurun_out/ncu_<family>.ncu-rep
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:gemm_tcgen05 -s 16 -c 8 -f -o gpurun_out/ncu_gemm python scripts/ncu_kernels.py gemm > gpurun_out/ncu_gemm.log 2>&1
timeout 600 $NCU -k regex:gn_ -s 20 -c 10 -f -o gpurun_out/ncu_gn python scripts/ncu_kernels.py gn > gpurun_out/ncu_gn.log 2>&1
timeout 600 $NCU -k regex:layernorm -s 4 -c 2 -f -o gpurun_out/ncu_ln python scripts/ncu_kernels.py ln > gpurun_out/ncu_ln.log 2>&1
timeout 600 $NCU -k regex:temporal_attn -s 4 -c 2 -f -o gpurun_out/ncu_attn python scripts/ncu_kernels.py attn > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_*.log
