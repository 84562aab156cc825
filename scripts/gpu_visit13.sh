#!/bin/bash
# New scenarios (nopositivity/eta, MFS threshold + radial + L-BFGS, flag-1 gradients), masks, normalize.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -s -k "scenario or mask or normalize or command_line" ) > gpurun_out/pytest_v13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v13.log
grep -n "^\[\|passed\|failed\|rc=\|^E  \|Error\|assert" gpurun_out/pytest_v13.log | tail -n 40
