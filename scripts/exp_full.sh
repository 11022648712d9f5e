#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 ) > gpurun_out/pytest_full.log 2>&1
cat gpurun_out/pytest_full.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
