#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "linear or layernorm_fold" 2>&1 | tail -4 )
CA_GEMM_NO_ASTAT=1 timeout 300 python scripts/gemm_lab.py noastat "m131072" > gpurun_out/gemm_lab_noastat.log 2>&1; cat gpurun_out/gemm_lab_noastat.log
timeout 300 python scripts/gemm_lab.py astat "m131072" > gpurun_out/gemm_lab_astat.log 2>&1; cat gpurun_out/gemm_lab_astat.log
CA_GEMM_TIMING=1 timeout 300 python scripts/gemm_lab.py t "m131072" 2>&1 | grep timing
