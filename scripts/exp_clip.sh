#!/bin/bash
# clip layout (windows + ControlNet servers) on N GPUs: parity check, then the config-3 bench
set -u
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/multirank_check.py clip > gpurun_out/clip_check_n$N.log 2>&1
echo "check exit $?"; tail -15 gpurun_out/clip_check_n$N.log
if [ "${2:-}" != "" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 4 --warmup 3 --workload config3 --parallelism clip $2 > gpurun_out/bench_clip_n$N.json 2> gpurun_out/bench_clip_n$N.err
echo "bench exit $?"; tail -5 gpurun_out/bench_clip_n$N.err; cut -c1-900 gpurun_out/bench_clip_n$N.json
fi
