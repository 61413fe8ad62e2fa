set -x
GK_TRACE=1 timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "stranger" -s 2>&1 | grep -E "trace\]|passed|failed" | grep -v "pool alloc" | head -20
