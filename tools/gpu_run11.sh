set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests11.log 2>&1; tail -3 gpurun_out/tests11.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench11.log 2>&1; tail -1 gpurun_out/bench11.log | cut -c1-300
bash tools/gpu_ncu_aux.sh r01aux3 "tie_fix|pack_keys"
