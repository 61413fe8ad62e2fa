set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests13.log 2>&1; tail -3 gpurun_out/tests13.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench13.log 2>&1; tail -1 gpurun_out/bench13.log | cut -c1-300
bash tools/gpu_multi.sh 2 r01q
