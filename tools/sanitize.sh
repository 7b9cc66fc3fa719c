#!/bin/bash
# compute-sanitizer passes over the small GPU tests (memcheck + racecheck); output in gpurun_out/sanitizer_*.log
set -u
mkdir -p gpurun_out
SEL='test_index_structures_edge_cases or test_linearize_blocks_match_oracle and scene_small or test_full_lm_solve_matches_oracle and scene_small or test_pnp_matches_oracle or test_dense_cholesky_solve and 258 or test_dense_cholesky_solve and 132 or test_visibility_noisy_poses or test_triangulation_options_edge_cases or test_total_optimisation_edge_cases or test_options_rejections_and_edge_cases or test_partitioned_band_solve or test_loop_closures_match_oracle'
for tool in ${SAN_TOOLS:-memcheck racecheck}; do
  timeout ${SAN_TIMEOUT:-300} compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py tests/test_front.py tests/test_calib.py tests/test_posegraph.py -m gpu -q -x -k "$SEL" \
    > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_$tool.log | tail -3
done
