#!/bin/bash
# First GPU bring-up: smoke, parity tests, a short bench of both arms, a launch list.
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --scale 0.1 --steps 3 --warmup 3 > gpurun_out/bench_scale0.1.json 2> gpurun_out/bench_scale0.1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 --ref-sample 100000 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -5 gpurun_out/smoke.log gpurun_out/pytest_gpu.log gpurun_out/bench_scale0.1.json gpurun_out/bench_ref.json
