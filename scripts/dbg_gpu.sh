python examples/closed_loop.py
python examples/closed_loop.py --no-graph
timeout 600 python -m pytest tests/test_ext_gpu.py -m gpu -q -x -k "closed_loop or graph or env_step or collision" 2>&1 | tail -3
