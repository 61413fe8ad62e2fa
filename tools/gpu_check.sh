# usage (under gpurun, one GPU): bash tools/gpu_check.sh <tag>
# the round-end sequence in one call: GPU tests, smoke(), the bench (both arms), launch list of one step
set -x
TAG=${1:-check}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests_$TAG.log 2>&1; tail -3 gpurun_out/tests_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/bench_$TAG.log 2>&1; tail -1 gpurun_out/bench_$TAG.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; tail -1 gpurun_out/bench_ref_$TAG.log | cut -c1-200
bash tools/gpu_launchlist.sh $TAG
