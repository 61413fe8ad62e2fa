# usage: bash tools/gpu_multi.sh <N> <tag>  -- bench at N GPUs (torchrun), plus the reference arm launch contract
set -x
N=${1:-2}
TAG=${2:-r01}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n$N.log 2>&1
tail -1 gpurun_out/bench_${TAG}_n$N.log | cut -c1-1500
