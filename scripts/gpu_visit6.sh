#!/bin/bash
mkdir -p gpurun_out
GVM_PROFILE_HOST=1 timeout 600 python bench.py --config c1 --steps 2 --warmup 3 --recon-iters 50 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1_hostprofile.txt; echo "c1 rc=$?"
tail -n 1 gpurun_out/bench_c1.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['recon'])"
head -n 40 gpurun_out/bench_c1_hostprofile.txt
