# usage (under gpurun, one GPU): bash tools/gpu_r2_quick.sh <tag>  -- GPU suite + bench runs
set -x
TAG=${1:-q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/tests_$TAG.log 2>&1; tail -12 gpurun_out/tests_$TAG.log
for i in 1 2; do
GK_TRACE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_$i.log 2> gpurun_out/bench_${TAG}_$i.err; grep "level1" gpurun_out/bench_${TAG}_$i.err | tail -1
tail -1 gpurun_out/bench_${TAG}_$i.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['step_wall_ms'], d['roofline']['stage_ms'], d['host_phase_ms_create_sort_count_destroy'])"
done
