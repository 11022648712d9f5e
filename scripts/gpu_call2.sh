#!/bin/bash
# gpurun call: GPU parity tests, GroupNorm ring v2 sweep, LayerNorm ring, ncu --set full of both, bench.
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-200
mb() { tag=$1; fam=$2; shift; shift; env "$@" timeout 300 python scripts/microbench.py --quick --only $fam --iters 10 --out gpurun_out/mb_${fam}_$tag.json 2>&1 | grep -E "bfhwc|layernorm" | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('$tag', r['shape'], r['us'], r['frac_hbm'])"; }
mb ring16 gn CA_GN_RING_KB=16
mb ring8 gn CA_GN_RING_KB=8 CA_GN_RING_STAGES=8
mb ring24 gn CA_GN_RING_KB=24
mb ring16f4 gn CA_GN_RING_KB=16 CA_GN_RING_FOLDERS=4
mb lnring ln CA_LN_RING=1
mb lnring32 ln CA_LN_RING_KB=32 CA_LN_RING_STAGES=3
mb lnring10 ln CA_LN_RING_KB=10 CA_LN_RING_STAGES=8
mb lnold ln CA_LN_RING=0
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:gn_ring -s 4 -c 2 -f -o gpurun_out/ncu_gnring2 python scripts/ncu_kernels.py gn > gpurun_out/ncu_gnring2.log 2>&1
timeout 400 $NCU -k regex:layernorm -s 4 -c 2 -f -o gpurun_out/ncu_ln2 python scripts/ncu_kernels.py ln > gpurun_out/ncu_ln2.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_call2.json 2> gpurun_out/bench_call2.err
cat gpurun_out/bench_call2.json | cut -c1-300
