#!/bin/bash
# Host-layer parity scenarios against the reference (each reference run in its own process), then the full GPU suite.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests/test_host_gpu.py -m gpu -q -s > gpurun_out/pytest_host_gpu.log 2>&1; echo "host rc=$?" >> gpurun_out/pytest_host_gpu.log
tail -n 60 gpurun_out/pytest_host_gpu.log
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_host_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "all rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 15 gpurun_out/pytest_gpu.log
