# usage: bash tools/gpu_profile.sh <tag>   (run under gpurun, one GPU)
# 1. plain bench run (must exit 0 without ncu)  2. ncu launch list of the same command
# 3. one `ncu --set full` capture of an onesweep pass of the timed step
set -x
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --clock-mode off"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
# 4 steps (3 warm-up + 1 timed) x 4 main-sort passes; the small refinement sorts also match the regex,
# so skip by launch count: capture launch #2 of the LAST step's main sort = the 4th step's second big pass
ncu --set full --clock-control none --import-source on -k regex:onesweep_kernel -s 85 -c 1 \
    -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/plain_$TAG.log | cut -c1-300; tail -2 gpurun_out/ncu_full_$TAG.log
