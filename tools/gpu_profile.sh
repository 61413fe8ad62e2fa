# usage: bash tools/gpu_profile.sh <tag>   (run under gpurun, one GPU)
# 1. plain bench run (must exit 0 without ncu)  2. ncu launch list of the same command
# 3. `ncu --set full` captures of the three HBM-streaming kernels of the timed step: one onesweep pass,
#    the pack kernel, the tie-repair/flags kernel
set -x
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-verify --clock-mode off"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
# 4 steps x 4 main-sort passes: skip to the last step's second pass
ncu --set full --clock-control none --import-source on -k regex:onesweep_kernel -s 13 -c 1 \
    -o gpurun_out/prof_${TAG}_onesweep $CMD > gpurun_out/ncu_full_${TAG}_onesweep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pack_keys_kernel -s 3 -c 1 \
    -o gpurun_out/prof_${TAG}_pack $CMD > gpurun_out/ncu_full_${TAG}_pack.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tie_fix_flags_kernel -s 3 -c 1 \
    -o gpurun_out/prof_${TAG}_tiefix $CMD > gpurun_out/ncu_full_${TAG}_tiefix.log 2>&1
tail -1 gpurun_out/plain_$TAG.log | cut -c1-200; tail -2 gpurun_out/ncu_full_${TAG}_tiefix.log
