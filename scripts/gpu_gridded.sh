#!/bin/bash
# Gridded-mode visit: FFT-gradient parity tests, gridded scenarios against the reference, a scaled C5 bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_host_gpu.py -m gpu -q -s -k "gridded" > gpurun_out/pytest_gridded.log 2>&1; echo "gridded rc=$?" >> gpurun_out/pytest_gridded.log
grep -n "^\[\|gridded gradient\|passed\|failed\|rc=\|Error\|error" gpurun_out/pytest_gridded.log | tail -n 12
timeout 1200 python bench.py --config c5 --scale ${1:-0.05} --steps 5 --warmup 3 --recon-iters 10 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
tail -n 1 gpurun_out/bench_c5.json; tail -n 5 gpurun_out/bench_c5.err
