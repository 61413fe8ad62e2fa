# usage: bash tools/gpu_launchlist_config.sh <tag> <config...>  -- ncu launch list of tools/bench_configs.py <config>
set -x
TAG=${1:-c4}; shift
mkdir -p gpurun_out
CMD="python tools/bench_configs.py $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
tail -2 gpurun_out/plain_$TAG.log | cut -c1-400
