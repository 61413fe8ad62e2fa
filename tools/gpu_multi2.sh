# usage (under gpurun --gpus 2): bash tools/gpu_multi2.sh <tag>  -- two-GPU parity + bench
set -x
TAG=${1:-m2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    tests/multi_gpu_check.py > gpurun_out/multi_check_${TAG}_n2.log 2>&1; grep -c " ok" gpurun_out/multi_check_${TAG}_n2.log; tail -3 gpurun_out/multi_check_${TAG}_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_${TAG}_n2.log 2>&1
tail -1 gpurun_out/bench_${TAG}_n2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['ms_per_step'],3), round(d['value'],2), d['verified'], d['phase_ms_rank0'], d['per_rank_local_sort'], d['exchange']['nvlink_gbs_per_gpu_outbound'], d['e2e']['value'])"
