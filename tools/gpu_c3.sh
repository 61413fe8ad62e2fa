# BASELINE.json configs[2]: human-sized 3.1 Gbp, both strands, k=31, key-range sharded over 8 B200
set -x
mkdir -p gpurun_out
free -g | head -2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --bases 387500000 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_c3_n8.log 2>&1
tail -3 gpurun_out/bench_c3_n8.log | cut -c1-400
nvidia-smi --query-gpu=memory.used --format=csv | head -3
