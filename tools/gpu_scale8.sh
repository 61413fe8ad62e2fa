# usage (under gpurun --gpus 8): bash tools/gpu_scale8.sh <tag>   -- multi-GPU parity at 8, weak scaling 2/4/8, C3 human-sized
set -x
TAG=${1:-s8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    tests/multi_gpu_check.py > gpurun_out/multi_check_${TAG}_n8.log 2>&1; tail -12 gpurun_out/multi_check_${TAG}_n8.log
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_${TAG}_n$N.log 2>&1
  tail -1 gpurun_out/bench_${TAG}_n$N.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['ms_per_step'],3), round(d['value'],2), d['verified'], d['phase_ms_rank0'], d['per_rank_local_sort']['total_ms'], d['exchange']['nvlink_gbs_per_gpu_outbound'], d['e2e']['value'])"
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 8 --config c3 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_${TAG}_c3_n8.log 2>&1
tail -1 gpurun_out/bench_${TAG}_c3_n8.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3', round(d['ms_per_step'],3), round(d['value'],2), d['verified'], d['phase_ms_rank0'], d['per_rank_local_sort']['total_ms'], d['exchange'])"
tail -3 gpurun_out/bench_${TAG}_c3_n8.log | cut -c1-600
