#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
BNV_DEBUG_DISABLE=4096 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_forced_wide.txt
for m in "" "--unfused" "--no-graph" "--no-graph --unfused"; do python examples/closed_loop.py $m 2>&1 | head -1; done | tee gpurun_out/closed_loop.txt
python scripts/phase_stamps.py 2>&1 | tail -2 | tee gpurun_out/phase_stamps.txt
for c in ${CONFIGS:-c1 c2 c3 c4}; do
  timeout 600 python bench.py --config $c --steps ${STEPS:-2000} --warmup 20 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  tail -2 gpurun_out/bench_$c.err; python scripts/bench_summary.py < gpurun_out/bench_$c.json
done
