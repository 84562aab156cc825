#!/bin/bash
# GPU parity tests (all, no -x) with full log under gpurun_out/.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 40 gpurun_out/pytest_gpu.log
