#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "upsample or concat" 2>&1 | tail -3 )
( timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullwidth.py -m gpu -q -x 2>&1 | tail -4 )
for i in 1 2; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-yardstick > gpurun_out/bench_copy$i.json 2> gpurun_out/bench_copy$i.err
echo "run $i: $(python -c "import json;d=json.load(open('gpurun_out/bench_copy$i.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], {k:(round(v['ms_total'],2), round(v['frac'],2)) for k,v in d['kernels'].items()})" 2>&1 | tail -1)"
done
