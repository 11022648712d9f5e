#!/bin/bash
# frame-window weak scaling of the headline bench on N GPUs (the driver's SCALE run), plus the multi-rank parity checks
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_windows_n$N.json 2> gpurun_out/bench_windows_n$N.err
echo "bench N=$N exit $?"; grep "^{" gpurun_out/bench_windows_n$N.json | cut -c1-260
if [ "${2:-}" = "tests" ]; then ( timeout 1200 python -m pytest tests/test_gpu_multirank.py -q -x 2>&1 | tail -4 ); fi
