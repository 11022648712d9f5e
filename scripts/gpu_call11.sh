#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log | cut -c1-250
mb() { tag=$1; fam=$2; shift; shift; env "$@" timeout 300 python scripts/microbench.py --quick --only $fam --iters 10 --out gpurun_out/mb_${fam}_$tag.json 2>&1 | grep -E "bfhwc|layernorm|temporal" | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('$tag', r['shape'], r['us'], r['frac_hbm'])"; }
mb slab gn CA_X=1
mb noslab gn CA_GN_SLAB=0
timeout 900 python scripts/microbench.py --iters 8 --only attn --out gpurun_out/mb_attn_full.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print(r['shape'], r['us'], r['frac_hbm'])"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_call11.json 2> gpurun_out/bench_call11.err
tail -3 gpurun_out/bench_call11.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_call11.json'))
print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'traffic', d['roofline']['traffic'])
for k,v in d['kernels'].items(): print(k, v['launches'], round(v['ms_total'],2), round(v['us_per_launch'],1), round(v['frac'],3))
PY
