#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tests/multirank_check.py clip > gpurun_out/clip_check_n8.log 2>&1
echo "check exit $?"; tail -3 gpurun_out/clip_check_n8.log
for ln in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2961$ln bench.py --gpus 8 --steps 8 --warmup 4 --workload config3 --parallelism clip --windows 5 --local-nets $ln --no-cpu-baseline > gpurun_out/bench_clip_n8_l$ln.json 2> gpurun_out/bench_clip_n8_l$ln.err
echo "bench local_nets=$ln exit $?"; tail -3 gpurun_out/bench_clip_n8_l$ln.err; cut -c1-330 gpurun_out/bench_clip_n8_l$ln.json
done
