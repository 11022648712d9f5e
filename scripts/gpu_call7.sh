#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python scripts/microbench.py --quick --only xattn --iters 10 --out gpurun_out/mb_xattn.json 2>&1 | cut -c1-200
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:cross_attn -s 4 -c 2 -f -o gpurun_out/ncu_xattn python scripts/ncu_kernels.py xattn > gpurun_out/ncu_xattn.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_call7.json 2> gpurun_out/bench_call7.err
tail -3 gpurun_out/bench_call7.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_call7.json'))
print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
for k,v in d['kernels'].items(): print(k, v['launches'], round(v['ms_total'],2), round(v['us_per_launch'],1), round(v['frac'],3))
PY
