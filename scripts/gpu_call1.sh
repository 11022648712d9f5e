#!/bin/bash
# gpurun call: GPU parity tests, GroupNorm ring parameter sweep, LN/attention microbench, ncu --set full of the HBM kernels.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
mb() { tag=$1; shift; env "$@" timeout 300 python scripts/microbench.py --quick --only gn --iters 10 --out gpurun_out/mb_gn_$tag.json 2>&1 | grep bfhwc | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('$tag', r['shape'], r['us'], r['frac_hbm'])"; }
mb team CA_GN_RING=0
mb ring16 CA_GN_RING_KB=16
mb ring8 CA_GN_RING_KB=8 CA_GN_RING_STAGES=8
mb ring8l3 CA_GN_RING_KB=8 CA_GN_RING_STAGES=8 CA_GN_RING_LAG=3
mb ring16l2 CA_GN_RING_KB=16 CA_GN_RING_LAG=2
mb ring24 CA_GN_RING_KB=24
timeout 300 python scripts/microbench.py --quick --only ln,attn --iters 10 --out gpurun_out/mb_lnattn.json 2>&1 | cut -c1-220
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:gn_ring -s 4 -c 2 -f -o gpurun_out/ncu_gnring python scripts/ncu_kernels.py gn > gpurun_out/ncu_gnring.log 2>&1
timeout 400 $NCU -k regex:layernorm -s 4 -c 2 -f -o gpurun_out/ncu_ln python scripts/ncu_kernels.py ln > gpurun_out/ncu_ln.log 2>&1
timeout 400 $NCU -k regex:temporal_attn -s 4 -c 2 -f -o gpurun_out/ncu_attn python scripts/ncu_kernels.py attn > gpurun_out/ncu_attn.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_call1.json 2> gpurun_out/bench_call1.err
cat gpurun_out/bench_call1.json | cut -c1-600
ls -la gpurun_out/*.ncu-rep
