# usage: bash tools/gpu_launchlist.sh <tag>  -- ncu launch list (per-kernel durations) of one bench step
set -x
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --clock-mode off"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
tail -1 gpurun_out/plain_$TAG.log | cut -c1-200
