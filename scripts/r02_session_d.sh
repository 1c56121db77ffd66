#!/bin/bash
# epilogue A/B: distributed merge (default) vs round-1 schedule (bit 8192); forced wide; all configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
BNV_DEBUG_DISABLE=4096 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_forced_wide.txt
BNV_DEBUG_DISABLE=8192 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r1_schedule.txt
echo "--- stamps: default / r1 schedule / K=131072"
python scripts/phase_stamps.py 2>&1 | tail -2 | tee gpurun_out/phase_stamps.txt
BNV_DEBUG_DISABLE=8192 python scripts/phase_stamps.py 2>&1 | tail -1
python scripts/phase_stamps.py 131072 2>&1 | tail -1
for c in ${CONFIGS:-c1 c2 c3 c4}; do
  timeout 600 python bench.py --config $c --steps ${STEPS:-1000} --warmup 20 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  tail -2 gpurun_out/bench_$c.err; python scripts/bench_summary.py < gpurun_out/bench_$c.json
done
echo "--- c1 with the r1 schedule"
BNV_DEBUG_DISABLE=8192 timeout 600 python bench.py --config c1 --steps 1000 --warmup 20 2>/dev/null | python scripts/bench_summary.py
