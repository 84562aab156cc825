#!/bin/bash
# gridding parity + timings after the start-table change, scaled C5 bench, ncu --set full of the dominant kernel on the bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_host_gpu.py -m gpu -q -k "gridd" > gpurun_out/pytest_gridded.log 2>&1; echo "gridded rc=$?" >> gpurun_out/pytest_gridded.log
tail -n 4 gpurun_out/pytest_gridded.log
timeout 1200 python bench.py --config c5 --scale ${1:-0.05} --steps 5 --warmup 3 --recon-iters 10 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
tail -n 1 gpurun_out/bench_c5.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['preprocessing'], d['recon'], d['roofline']['achieved'])"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_grad_umma -s 3 -c 1 -f -o gpurun_out/r1b_umma_c2 \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --recon-iters 0 > gpurun_out/ncu_full_c2.log 2>&1; echo "ncu rc=$?"
tail -n 3 gpurun_out/ncu_full_c2.log; ls -la gpurun_out/*.ncu-rep
