set -x
for env in "GK_FRAGMENTS=1" "GK_FRAGMENTS=0" "GK_PEER_EXCHANGE=0" "GK_SORT_HYBRID=0"; do
echo "=== $env"
env $env timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/shard_verify.py --same-gpu --bases 50000000 2>&1 | grep -E "rank [01]:" | cut -c1-420
done
