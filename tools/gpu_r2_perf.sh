# usage (under gpurun, one GPU): bash tools/gpu_r2_perf.sh <tag> [ENV=VAL ...]   -- GPU suite, then the bench per env setting
set -x
TAG=${1:-p}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/tests_$TAG.log 2>&1; tail -6 gpurun_out/tests_$TAG.log
for env in "A=1" "$@"; do
  echo "=== $env"
  env $env timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --no-verify --clock-mode off > gpurun_out/bench_${TAG}.log 2> gpurun_out/bench_${TAG}.err
  tail -1 gpurun_out/bench_${TAG}.log | python -c "import sys,json,statistics; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'median wall', round(statistics.median(d['step_wall_ms']),3), {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()}, d['host_phase_ms_create_sort_count_destroy'][-1], d['gpu_launches'])"
done
