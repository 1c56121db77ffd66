#!/bin/bash
# GPU session: full parity suite, memcheck of the new kernel modes on small cases, bench line, widened-path rates, phase stamps.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ext_gpu.py -q -x -k "golden or batched_matches or differs" 2>&1 | tail -15 | tee gpurun_out/memcheck.txt
python bench.py --steps 5000 --warmup 50 > gpurun_out/bench_native.json 2> gpurun_out/bench_native.err; cat gpurun_out/bench_native.json; tail -3 gpurun_out/bench_native.err
timeout 600 python scripts/bench_configs.py 2>&1 | tee gpurun_out/bench_configs.jsonl | tail -20
python scripts/phase_stamps.py > gpurun_out/phase_stamps.txt 2>&1; tail -3 gpurun_out/phase_stamps.txt
