#!/bin/bash
# Re-validation visit: full GPU test suite (timed), smoke, default bench of both arms.
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 30 gpurun_out/pytest_gpu.log
( time timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; tail -n 5 gpurun_out/smoke.log
( time timeout 600 python bench.py --impl reference ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -n 1 gpurun_out/bench_ref.json
( time timeout 900 python bench.py ) > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"; tail -n 1 gpurun_out/bench_c2.json; tail -n 4 gpurun_out/bench_c2.err
