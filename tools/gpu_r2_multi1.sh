# usage (under gpurun, ONE GPU): bash tools/gpu_r2_multi1.sh <tag>  -- full GPU suite (includes two ranks on one GPU)
set -x
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/tests_multi_$TAG.log 2>&1; tail -30 gpurun_out/tests_multi_$TAG.log
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/tests_$TAG.log 2>&1; tail -12 gpurun_out/tests_$TAG.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.log 2> gpurun_out/bench_${TAG}.err
tail -1 gpurun_out/bench_${TAG}.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['verified'], d['verification'].get('report'), d['step_wall_ms'], d['roofline']['stage_ms'], d['e2e']['ms_per_step'])"
