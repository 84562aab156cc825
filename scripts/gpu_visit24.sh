#!/bin/bash
# compute-sanitizer: memcheck over (almost) the whole parity suite, racecheck + initcheck over the shared-memory kernels.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 97 --print-limit 20 \
  python -m pytest tests/test_parity_gpu.py tests/test_edges_gpu.py tests/test_cli_gpu.py -m gpu -q -x -k "not full_size" \
  > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -n "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitizer_memcheck.log | tail -n 4
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 97 --print-limit 20 \
  python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "gridding_bit_exact and Gaussian2D or weights_bit_exact and Briggs or conv or priors and Entropy or vector_ops or chi2_and_residuals or gradient_vs_fp64_oracle and 0-2" \
  > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -n "RACECHECK SUMMARY\|ERROR SUMMARY\|passed\|failed\|hazard" gpurun_out/sanitizer_racecheck.log | tail -n 6
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 97 --print-limit 20 \
  python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "umma_multitile and True or error_maps_vs or gridded_gradient or half_plane" \
  > gpurun_out/sanitizer_initcheck.log 2>&1; echo "initcheck rc=$?"
grep -n "ERROR SUMMARY\|passed\|failed\|Uninitialized" gpurun_out/sanitizer_initcheck.log | tail -n 6
