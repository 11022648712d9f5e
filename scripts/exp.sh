( timeout 1500 python -m pytest tests/test_gpu_fullwidth.py -q -x -s ) > gpurun_out/pytest_fullwidth.log 2>&1
grep -n "full width\|fp16\|lcm\|config 3\|passed\|failed\|Error" gpurun_out/pytest_fullwidth.log | head
timeout 900 python bench.py --workload config3 --steps 5 --warmup 3 > gpurun_out/bench_config3.json 2> gpurun_out/bench_config3.err; echo "bench c3 exit $?"; cut -c1-300 gpurun_out/bench_config3.json; tail -3 gpurun_out/bench_config3.err
