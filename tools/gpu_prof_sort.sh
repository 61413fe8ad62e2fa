set -x
TAG=${1:-r01b}
CFG=${2:-0}
export GK_SORT_CFG=$CFG
CMD="python tools/bench_sort.py --child 200000000"
$CMD > gpurun_out/sortplain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:onesweep -s 20 -c 1 \
    -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_sort_$TAG.log 2>&1
cat gpurun_out/sortplain_$TAG.log; tail -2 gpurun_out/ncu_sort_$TAG.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not config2" > gpurun_out/tests3.log 2>&1; tail -5 gpurun_out/tests3.log
