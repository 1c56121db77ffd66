#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
BNV_DEBUG_DISABLE=4096 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_forced_wide.txt
python scripts/phase_stamps.py 2>&1 | tail -2 | tee gpurun_out/phase_stamps.txt
python scripts/phase_stamps.py 131072 2>&1 | tail -1 | tee -a gpurun_out/phase_stamps.txt
for c in ${CONFIGS:-c1 c2 c3 c4}; do
  timeout 600 python bench.py --config $c --steps ${STEPS:-2000} --warmup 20 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  tail -2 gpurun_out/bench_$c.err; python scripts/bench_summary.py < gpurun_out/bench_$c.json
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_ext_gpu.py tests/test_modes_gpu.py -m gpu -q -x -k "prelaunched or golden_single or stochastic_golden or batched_matches or env_step or collision or dwa or graph or top_samples or lean_solver_golden or general_angle or setters_cancel or closed_loop_example" 2>&1 | tail -4 | tee gpurun_out/sanitize_memcheck.txt
BNV_DEBUG_DISABLE=4096 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_ext_gpu.py -m gpu -q -x -k "golden_single or stochastic_golden or batched_matches or top_samples" 2>&1 | tail -4 | tee gpurun_out/sanitize_memcheck_wide.txt
