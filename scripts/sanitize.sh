#!/bin/bash
# compute-sanitizer passes over small parity cases of every kernel mode (memcheck, racecheck, synccheck)
mkdir -p gpurun_out
SEL='prelaunched or golden_single or stochastic_golden or batched_matches or env_step or collision or risk_map_golden or dwa or without_a_staged or graph or top_samples or lean_solver_golden or lean_solver_full_size or general_angle or setters_cancel or closed_loop_example or two_stage'
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_ext_gpu.py tests/test_modes_gpu.py -m gpu -q -x -k "$SEL" 2>&1 | tail -6 | tee gpurun_out/sanitize_$tool.txt
  # the same golden cases through the wide variant of the rollout kernel (chunked flushes, two-level merge, normalize kernel)
  BNV_DEBUG_DISABLE=4096 timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_ext_gpu.py -m gpu -q -x -k "golden_single or stochastic_golden or batched_matches or top_samples" 2>&1 | tail -4 | tee gpurun_out/sanitize_${tool}_wide.txt
done
