#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
BNV_DEBUG_DISABLE=4096 timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_forced_wide.txt
python scripts/bench_lean.py 2>&1 | tee gpurun_out/bench_lean.jsonl
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_modes_gpu.py -m gpu -q -x -k "top_samples or lean_solver_golden" 2>&1 | tail -4
