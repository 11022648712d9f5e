#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -k "groupnorm or smoke or unet or model" ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log | cut -c1-250
mb() { tag=$1; fam=$2; shift; shift; env "$@" timeout 300 python scripts/microbench.py --quick --only $fam --iters 10 --out gpurun_out/mb_${fam}_$tag.json 2>&1 | grep -E "bfhwc|layernorm|temporal" | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('$tag', r['shape'], r['us'], r['frac_hbm'])"; }
mb cluster8 gn CA_X=1
mb cluster4 gn CA_GN_SLAB_CLUSTER=4
mb cluster1 gn CA_GN_SLAB_CLUSTER=1
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:gn_slab -s 4 -c 2 -f -o gpurun_out/ncu_gnslab python scripts/ncu_kernels.py gn > gpurun_out/ncu_gnslab.log 2>&1
