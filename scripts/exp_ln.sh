#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "layernorm_fold or linear" 2>&1 | tail -5 ) > gpurun_out/t_ln.log 2>&1
( timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullwidth.py tests/test_gpu_dropin.py -m gpu -q -x -s 2>&1 | tail -25 ) > gpurun_out/t_model.log 2>&1
timeout 60 ./scripts/ubench/softmax_pipes > gpurun_out/softmax_pipes.txt 2>&1
for cfg in "0 0" "1 0" "0 1" "1 1" "1 0" "0 0"; do
  set -- $cfg
  CA_LN_FOLD=$1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-yardstick --stream-overlap $2 > gpurun_out/bench_fold$1_ov$2.json 2> gpurun_out/bench_fold$1_ov$2.err
  echo "fold=$1 overlap=$2: $(python -c "import json;d=json.load(open('gpurun_out/bench_fold$1_ov$2.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], {k:round(v['ms_total'],2) for k,v in d['kernels'].items()})" 2>&1 | tail -1)"
done
cat gpurun_out/t_ln.log gpurun_out/t_model.log; cat gpurun_out/softmax_pipes.txt
