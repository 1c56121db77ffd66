#!/bin/bash
# Round-2 single-GPU session: parity suite (default kernel selection, then every eligible solver forced onto the wide
# variant), bench lines of every configuration.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
BNV_DEBUG_DISABLE=4096 timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_forced_wide.txt
for c in ${CONFIGS:-c1 c2 c3 c4}; do
  timeout 600 python bench.py --config $c --steps ${STEPS:-1000} --warmup 20 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  tail -2 gpurun_out/bench_$c.err; python scripts/bench_summary.py < gpurun_out/bench_$c.json
done
BNV_DEBUG_DISABLE=4096 timeout 600 python bench.py --config c1 --steps 1000 --warmup 20 2>/dev/null | python scripts/bench_summary.py
