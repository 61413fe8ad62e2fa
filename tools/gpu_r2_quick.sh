# usage (under gpurun, one GPU): bash tools/gpu_r2_quick.sh <tag>  -- GPU suite + one traced bench run
set -x
TAG=${1:-q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/tests_$TAG.log 2>&1; tail -12 gpurun_out/tests_$TAG.log
GK_TRACE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; grep "level1" gpurun_out/bench_$TAG.err | tail -2; tail -1 gpurun_out/bench_$TAG.log | cut -c1-2500
