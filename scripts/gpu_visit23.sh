#!/bin/bash
# compute-sanitizer memcheck over the small parity tests (every kernel family once).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 97 --print-limit 20 --launch-timeout 120 \
  python -m pytest tests/test_parity_gpu.py tests/test_edges_gpu.py -m gpu -q -x \
  -k "not full_size and not multitile and not wterm_exact and (ragged_visibility_counts and 33 or empty_block or zero_weights or off_the_tile and 96 or mosaic or error_maps_vs or error_maps_gridded or half_plane or gridding_bit_exact and Gaussian2D or weights_bit_exact and Briggs or priors and Entropy or vector_ops or chi2_and_residuals)" \
  > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -n "ERROR SUMMARY\|Invalid\|passed\|failed\|=========     at\|Saved host" gpurun_out/sanitizer_memcheck.log | head -n 30
tail -n 5 gpurun_out/sanitizer_memcheck.log
