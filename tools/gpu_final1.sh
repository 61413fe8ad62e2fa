# usage (under gpurun, one GPU): bash tools/gpu_final1.sh <tag>  -- what the driver runs at round end, plus the other workloads
set -x
TAG=${1:-final}
mkdir -p gpurun_out
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log | cut -c1-200
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.log 2> gpurun_out/bench_${TAG}_reference.err; tail -1 gpurun_out/bench_${TAG}_reference.log | cut -c1-900
timeout 900 python bench.py > gpurun_out/bench_${TAG}_n1.log 2> gpurun_out/bench_${TAG}_n1.err; tail -1 gpurun_out/bench_${TAG}_n1.log | cut -c1-3000
timeout 600 python bench.py --workload repeats --no-cpu-baseline > gpurun_out/bench_${TAG}_repeats.log 2> gpurun_out/bench_${TAG}_repeats.err
tail -1 gpurun_out/bench_${TAG}_repeats.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('repeats', d['ms_per_step'], d['value'], d['verified'], d['roofline']['stage_ms'], d['e2e']['ms_per_step'])"
timeout 900 python tools/bench_configs.py > gpurun_out/other_configs_$TAG.jsonl 2> gpurun_out/other_configs_$TAG.err; tail -12 gpurun_out/other_configs_$TAG.jsonl | cut -c1-260
