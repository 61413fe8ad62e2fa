set -x
mkdir -p gpurun_out
timeout 600 python tools/bench_sort.py 200000000 > gpurun_out/sortcfg.log 2>&1
cat gpurun_out/sortcfg.log
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -x -q -m gpu -k "not config2" > gpurun_out/tests2.log 2>&1; tail -5 gpurun_out/tests2.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.log 2>&1; tail -2 gpurun_out/bench2.log
