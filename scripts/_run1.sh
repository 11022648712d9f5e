set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "hoisted or fused_loop or cuda_graph or denoising_step or own_streams" 2>&1 | tail -8 > gpurun_out/t_loop.log; cat gpurun_out/t_loop.log
for h in 0 1 0 1; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-yardstick --hoist-cond-embedding $h 2> gpurun_out/bench_hoist$h.err > gpurun_out/bench_hoist$h.json; python -c "
import json;d=json.load(open('gpurun_out/bench_hoist$h.json'));print('hoist $h', d['ms_per_step'], d['e2e']['ms_per_step'], d['cond_embedding'], d['gpu_launches'])"
done
