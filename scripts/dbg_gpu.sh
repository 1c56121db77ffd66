python examples/closed_loop.py
timeout 600 python -m pytest tests/test_ext_gpu.py -m gpu -q -x -k "closed_loop" 2>&1 | tail -3
