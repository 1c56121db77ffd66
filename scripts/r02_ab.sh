#!/bin/bash
for rep in 1 2; do
for mode in 0 8192; do
  echo -n "BNV_DEBUG_DISABLE=$mode: "; BNV_DEBUG_DISABLE=$mode timeout 600 python bench.py --config c1 --steps 3000 --warmup 50 2>/dev/null | python scripts/bench_summary.py | cut -c1-200
done; done
BNV_DEBUG_DISABLE=8192 python scripts/phase_stamps.py 2>&1 | tail -2
