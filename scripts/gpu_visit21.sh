#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_host_gpu.py -m gpu -q -s -k "offset_field or cg_natural" ) > gpurun_out/pytest_v21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v21.log
grep -n "^\[\|passed\|failed\|rc=\|^E  \|Error" gpurun_out/pytest_v21.log | tail -n 20
