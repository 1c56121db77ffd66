timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo split on; python scripts/bench_bigk.py 2>&1 | cut -c1-200
echo split off; BNV_DEBUG_DISABLE=1024 python scripts/bench_bigk.py 2>&1 | cut -c1-200
