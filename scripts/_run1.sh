set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/pytest_verify.log 2>&1
cat gpurun_out/pytest_verify.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py 2> gpurun_out/bench_final.err > gpurun_out/bench_final.json; cut -c1-400 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
