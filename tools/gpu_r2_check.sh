# usage (under gpurun, one GPU): bash tools/gpu_r2_check.sh <tag>
# round-2 check: new tests first, full GPU suite, smoke, bench (fragments on / off), launch list of one step
set -x
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -m gpu > gpurun_out/tests_new_$TAG.log 2>&1; tail -15 gpurun_out/tests_new_$TAG.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/tests_$TAG.log 2>&1; tail -15 gpurun_out/tests_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2>&1; tail -1 gpurun_out/bench_$TAG.log | cut -c1-1500
GK_FRAGMENTS=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_nofrag.log 2>&1; tail -1 gpurun_out/bench_${TAG}_nofrag.log | cut -c1-400
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --clock-mode off"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
python tools/show_launches.py gpurun_out/launches_$TAG.csv 60 | tail -70
