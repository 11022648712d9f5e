#!/bin/bash
# final verification of a tree: the whole GPU suite, the smoke entry and a short bench
set -u
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/pytest_verify.log 2>&1
cat gpurun_out/pytest_verify.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-yardstick 2> gpurun_out/bench_verify.err | cut -c1-400
