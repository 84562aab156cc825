#!/bin/bash
# Whole GPU suite after the FITS / half-plane / mask changes (timed), smoke.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -n "passed\|failed\|rc=\|^E  \|Error\|^real\|s call" gpurun_out/pytest_gpu.log | tail -n 24
( time timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; tail -n 4 gpurun_out/smoke.log
