#!/bin/bash
# One GPU visit: host-layer scenario parity (only the Briggs one by default), a short C1 bench, the C2 bench of both arms.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_gpu.py -m gpu -q -s -k "${1:-briggs}" > gpurun_out/pytest_host_gpu.log 2>&1; echo "host rc=$?" >> gpurun_out/pytest_host_gpu.log
grep -n "^\[\|passed\|failed\|rc=" gpurun_out/pytest_host_gpu.log | tail
timeout 600 python bench.py --config c1 --steps 5 --warmup 3 --recon-iters 50 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "c1 rc=$?"
tail -n 2 gpurun_out/bench_c1.json; tail -n 3 gpurun_out/bench_c1.err
bash scripts/gpu_bench.sh
