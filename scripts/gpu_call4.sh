#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-200
mb() { tag=$1; fam=$2; shift; shift; env "$@" timeout 300 python scripts/microbench.py --quick --only $fam --iters 10 --out gpurun_out/mb_${fam}_$tag.json 2>&1 | grep -E "bfhwc|layernorm|temporal" | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('$tag', r['shape'], r['us'], r['frac_hbm'])"; }
mb c3k16 gn CA_GN_RING_CTAS=3 CA_GN_RING_KB=16
mb c3k12 gn CA_GN_RING_CTAS=3 CA_GN_RING_KB=12
mb c3k20 gn CA_GN_RING_CTAS=3 CA_GN_RING_KB=20
mb c2k24 gn CA_GN_RING_CTAS=2 CA_GN_RING_KB=24
mb c2k32 gn CA_GN_RING_CTAS=2 CA_GN_RING_KB=32
mb c2k16 gn CA_GN_RING_CTAS=2 CA_GN_RING_KB=16 CA_GN_RING_STAGES=6
mb lnattn ln,attn CA_X=1
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:gn_ring -s 4 -c 2 -f -o gpurun_out/ncu_gnring4 python scripts/ncu_kernels.py gn > gpurun_out/ncu_gnring4.log 2>&1
timeout 400 $NCU -k regex:temporal_attn -s 4 -c 2 -f -o gpurun_out/ncu_attn4 python scripts/ncu_kernels.py attn > gpurun_out/ncu_attn4.log 2>&1
timeout 400 $NCU -k regex:layernorm -s 4 -c 2 -f -o gpurun_out/ncu_ln4 python scripts/ncu_kernels.py ln > gpurun_out/ncu_ln4.log 2>&1
