#!/bin/bash
# One gpurun call: GPU parity tests, smoke, headline bench (+ reference arm), microbench, ncu launch list (time + DRAM bytes per
# launch) of one eager denoising step.  Everything lands under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 1800 python -m pytest tests -m gpu -q -s ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1200 python bench.py --steps 5 --warmup 3 --torch-profile gpurun_out/torch_profile.txt > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 900 python scripts/microbench.py --quick --out gpurun_out/microbench_quick.json > gpurun_out/microbench.log 2>&1
CA_NCU_RANGE=1 timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
   --profile-from-start off -c 20000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-eager-yardstick > gpurun_out/bench_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/bench_ncu.log
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_reference.json
