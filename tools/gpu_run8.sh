set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests8.log 2>&1; tail -5 gpurun_out/tests8.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench8.log 2>&1; tail -1 gpurun_out/bench8.log | cut -c1-300
GK_SORT_HYBRID=0 timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench8_plain.log 2>&1; tail -1 gpurun_out/bench8_plain.log | cut -c1-300
