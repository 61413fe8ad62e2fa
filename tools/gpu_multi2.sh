# usage (under gpurun --gpus 2): bash tools/gpu_multi2.sh <tag>   -- multi-GPU parity (2 GPUs + 2 ranks on 1 GPU), bench N=2
set -x
TAG=${1:-m2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/tests_multi_$TAG.log 2>&1; tail -25 gpurun_out/tests_multi_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n2.log 2>&1
tail -1 gpurun_out/bench_${TAG}_n2.log | cut -c1-3000
