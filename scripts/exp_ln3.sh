#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "layernorm_fold or linear" 2>&1 | tail -3 )
timeout 600 python scripts/ln_fold_lab.py > gpurun_out/ln_fold_lab.log 2>&1
cat gpurun_out/ln_fold_lab.log
for cfg in 0 1 0 1; do
  CA_LN_FOLD=$cfg timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-yardstick > gpurun_out/bench_fold$cfg.json 2> gpurun_out/bench_fold$cfg.err
  echo "fold=$cfg: $(python -c "import json;d=json.load(open('gpurun_out/bench_fold$cfg.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'])" 2>&1 | tail -1)"
done
