set -x
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "both_strand_layout or longer_than" 2>&1 | tail -5
timeout 800 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -30
