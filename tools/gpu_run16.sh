set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench16.log 2>&1; tail -1 gpurun_out/bench16.log | cut -c1-300
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench16_ref.log 2>&1; tail -1 gpurun_out/bench16_ref.log | cut -c1-300
timeout 1200 python tools/bench_configs.py c1 c4 c5 > gpurun_out/configs16.log 2>&1; cat gpurun_out/configs16.log | cut -c1-330
