#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --config c1 --steps 5 --warmup 3 --recon-iters 50 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "c1 rc=$?"
tail -n 1 gpurun_out/bench_c1.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['recon'])"
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_host_gpu.py -m gpu -q -k "gridd" > gpurun_out/pytest_gridded.log 2>&1; tail -n 3 gpurun_out/pytest_gridded.log
timeout 1200 python bench.py --config c5 --scale 0.05 --steps 5 --warmup 3 --recon-iters 10 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
tail -n 1 gpurun_out/bench_c5.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['preprocessing'], d['recon'])"
