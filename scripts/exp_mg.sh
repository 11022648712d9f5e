set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/mg_gpus.txt
( timeout 900 python -m pytest tests/test_gpu_multirank.py -q -x ) > gpurun_out/mg_pytest.log 2>&1
tail -15 gpurun_out/mg_pytest.log
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --parallelism cfg > gpurun_out/bench_n2_cfg.json 2> gpurun_out/bench_n2_cfg.err; echo "cfg exit $?"; cut -c1-400 gpurun_out/bench_n2_cfg.json; tail -3 gpurun_out/bench_n2_cfg.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 --parallelism controlnet > gpurun_out/bench_n2_cn.json 2> gpurun_out/bench_n2_cn.err; echo "cn exit $?"; cut -c1-400 gpurun_out/bench_n2_cn.json; tail -3 gpurun_out/bench_n2_cn.err
fi
if [ "$NG" -ge 4 ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 4 --steps 5 --warmup 3 --parallelism cfg+controlnet > gpurun_out/bench_n4_cfgcn.json 2> gpurun_out/bench_n4_cfgcn.err; echo "cfg+cn exit $?"; cut -c1-400 gpurun_out/bench_n4_cfgcn.json; tail -3 gpurun_out/bench_n4_cfgcn.err
fi
