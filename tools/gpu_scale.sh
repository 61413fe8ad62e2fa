# usage: bash tools/gpu_scale.sh <tag> <N...>   -- parity check at the largest N, then the bench at every N
set -x
TAG=$1; shift
mkdir -p gpurun_out
LAST=${@: -1}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $LAST --master-addr 127.0.0.1 --master-port 29541 \
    tests/multi_gpu_check.py > gpurun_out/multi_check_${TAG}_n$LAST.log 2>&1; tail -8 gpurun_out/multi_check_${TAG}_n$LAST.log
for N in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_n1.log 2>&1
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n$N.log 2>&1
  fi
  tail -1 gpurun_out/bench_${TAG}_n$N.log | cut -c1-260
done
