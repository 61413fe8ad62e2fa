set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests9.log 2>&1; tail -5 gpurun_out/tests9.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench9.log 2>&1; tail -1 gpurun_out/bench9.log | cut -c1-300
bash tools/gpu_launchlist.sh r01h
