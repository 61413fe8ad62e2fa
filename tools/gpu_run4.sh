set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.log 2>&1; tail -1 gpurun_out/bench4.log | cut -c1-3000
export GK_SORT_CFG=0
CMD="python tools/bench_sort.py --child 200000000"
$CMD > gpurun_out/sortplain_r01c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:onesweep -s 20 -c 1 \
    -o gpurun_out/prof_r01c $CMD > gpurun_out/ncu_sort_r01c.log 2>&1
tail -2 gpurun_out/ncu_sort_r01c.log
