( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
LAB_CUBLAS=1 timeout 200 python scripts/gemm_lab.py final > gpurun_out/lab_final.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
tail -5 gpurun_out/pytest_gpu.log; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
