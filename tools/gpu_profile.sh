# usage: bash tools/gpu_profile.sh <tag>   (run under gpurun; one ncu tool per call)
set -x
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:onesweep -s 72 -c 2 \
    -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/plain_$TAG.log gpurun_out/ncu_list_$TAG.log gpurun_out/ncu_full_$TAG.log
