#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/ln_fold_lab.py > gpurun_out/ln_fold_lab.log 2>&1
cat gpurun_out/ln_fold_lab.log
( timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "streams" 2>&1 | tail -5 )
for cfg in "1 0" "1 1" "1 0" "1 1"; do
  set -- $cfg
  CA_LN_FOLD=$1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-yardstick --stream-overlap $2 > gpurun_out/bench_fold$1_ov$2.json 2> gpurun_out/bench_fold$1_ov$2.err
  echo "fold=$1 overlap=$2: $(python -c "import json;d=json.load(open('gpurun_out/bench_fold$1_ov$2.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'])" 2>&1 | tail -1)"
  tail -3 gpurun_out/bench_fold$1_ov$2.err
done
