timeout 600 python -m pytest tests/test_ext_gpu.py -m gpu -q -x -k "without_a_staged or batch_of_one" 2>&1 | tail -12
