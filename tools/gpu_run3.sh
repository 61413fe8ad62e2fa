set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/tests_k.log 2>&1; tail -3 gpurun_out/tests_k.log
timeout 600 python tools/bench_sort.py 200000000 > gpurun_out/sortcfg3.log 2>&1; cat gpurun_out/sortcfg3.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not config2" > gpurun_out/tests_p.log 2>&1; tail -3 gpurun_out/tests_p.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench3.log 2>&1; tail -1 gpurun_out/bench3.log
