# full validation of the current tree on one B200: tests, smoke, bench (both arms), ncu launch list
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests6.log 2>&1; tail -3 gpurun_out/tests6.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke6.log 2>&1; tail -1 gpurun_out/smoke6.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench6.log 2>&1; tail -1 gpurun_out/bench6.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench6_ref.log 2>&1; tail -1 gpurun_out/bench6_ref.log | cut -c1-600
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_r01f.csv $CMD > gpurun_out/ncu_list_r01f.log 2>&1
tail -2 gpurun_out/ncu_list_r01f.log | cut -c1-300
