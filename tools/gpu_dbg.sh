timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -m gpu -k "64bit or longer_than_one_key_word or long_k or k64 or C4" 2>&1 | tail -12
timeout 800 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -12
