# usage: bash tools/gpu_ncu_aux.sh <tag> <regex>  -- full ncu capture of the non-sort kernels of one bench step
set -x
TAG=${1:-r01aux}
RE=${2:-"tie_fix|flag_group_hist|pack_keys|digit_hist|pack4_words"}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --clock-mode off"
ncu --set full --clock-control none --import-source on -k "regex:$RE" -c 5 \
    -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
