set -x
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "longer_than_one_key_word" 2>&1 | tail -5
timeout 800 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -30
GK_FRAG_PRESORT=0 timeout 800 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/owner_like.py 0.35 2>&1 | tail -4 | cut -c1-900
