#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -s -k "nopositivity or normalize" ) > gpurun_out/pytest_v14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v14.log
grep -n "^\[\|passed\|failed\|rc=\|^E  \|Error" gpurun_out/pytest_v14.log | tail -n 20
