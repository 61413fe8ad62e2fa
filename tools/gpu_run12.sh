set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/tests12.log 2>&1; tail -15 gpurun_out/tests12.log
bash tools/gpu_multi.sh 2 r01p
GK_PEER_EXCHANGE=0 bash tools/gpu_multi.sh 2 r01p_nccl
