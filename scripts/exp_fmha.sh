#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "spatial_attention" 2>&1 | grep -v "^$" | tail -30 ) > gpurun_out/fmha_test.log 2>&1
grep -n "timed out\|rror\|assert\|FAILED\|passed\|failed" gpurun_out/fmha_test.log | head -8
timeout 200 python scripts/fmha_lab.py 2>&1 | tail -4
CA_FMHA_TIMING=1 timeout 200 python scripts/fmha_lab.py 2>&1 | grep timing | head -1
