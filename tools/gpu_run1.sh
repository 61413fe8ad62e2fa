set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not config2" > gpurun_out/parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "config2" > gpurun_out/parity_c2.log 2>&1; echo "c2 rc=$?" >> gpurun_out/summary.txt
tail -5 gpurun_out/kernels.log gpurun_out/parity.log gpurun_out/smoke.log gpurun_out/bench.log gpurun_out/parity_c2.log
cat gpurun_out/summary.txt
