#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -k "top_samples or lean or dwa or closed_loop" 2>&1 | tail -3
python scripts/bench_lean.py 2>&1 | cut -c1-120
