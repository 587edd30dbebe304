timeout 600 python -m pytest tests/test_gpu_own_fft.py tests/test_gpu_round2.py -x -q -m gpu 2>&1 | tail -4
