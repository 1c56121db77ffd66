#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
BNV_DEBUG_DISABLE=4096 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_forced_wide.txt
python scripts/phase_stamps.py 131072 2>&1 | tail -1 | tee gpurun_out/phase_stamps_k131072.txt
for c in ${CONFIGS:-c2 c1}; do
  timeout 600 python bench.py --config $c --steps ${STEPS:-2000} --warmup 20 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  tail -2 gpurun_out/bench_$c.err; python scripts/bench_summary.py < gpurun_out/bench_$c.json
done
BNV_DEBUG_DISABLE=4096 timeout 600 python bench.py --config c4 --steps 1000 --warmup 20 2>/dev/null | python scripts/bench_summary.py
BNV_DEBUG_DISABLE=4096 timeout 600 python bench.py --config c3 --steps 1000 --warmup 20 2>/dev/null | python scripts/bench_summary.py
