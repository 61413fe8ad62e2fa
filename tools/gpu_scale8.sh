set -x
mkdir -p gpurun_out
TAG=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n8.log 2>&1
tail -1 gpurun_out/bench_${TAG}_n8.log | cut -c1-260
