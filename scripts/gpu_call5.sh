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
mb dflt gn,ln CA_X=1
mb k32s3 gn CA_GN_RING_KB=32 CA_GN_RING_STAGES=3
mb k40 gn CA_GN_RING_KB=40
mb f4 gn CA_GN_RING_FOLDERS=4
mb ln16s5 ln CA_LN_RING_KB=16 CA_LN_RING_STAGES=5
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_call5.json 2> gpurun_out/bench_call5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_call5.json'))
print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
for k,v in d['kernels'].items(): print(k, v['launches'], round(v['ms_total'],2), round(v['us_per_launch'],1), round(v['frac'],3))
PY
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:gn_ring -s 4 -c 2 -f -o gpurun_out/ncu_gnring5 python scripts/ncu_kernels.py gn > gpurun_out/ncu_gnring5.log 2>&1
timeout 400 $NCU -k regex:layernorm -s 4 -c 2 -f -o gpurun_out/ncu_ln5 python scripts/ncu_kernels.py ln > gpurun_out/ncu_ln5.log 2>&1
