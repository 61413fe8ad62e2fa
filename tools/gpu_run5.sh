set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "radix or partition" > gpurun_out/tests_k5.log 2>&1; tail -3 gpurun_out/tests_k5.log
timeout 600 python tools/bench_sort.py 200000000 0 2 3 > gpurun_out/sortcfg5.log 2>&1; cat gpurun_out/sortcfg5.log
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.log 2>&1; tail -1 gpurun_out/bench5.log | cut -c1-200
export GK_SORT_CFG=0
CMD="python tools/bench_sort.py --child 200000000"
$CMD > gpurun_out/sortplain_r01d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:onesweep -s 20 -c 1 \
    -o gpurun_out/prof_r01d $CMD > gpurun_out/ncu_sort_r01d.log 2>&1
tail -2 gpurun_out/ncu_sort_r01d.log
