#!/bin/bash
# Bench both arms + ncu launch list of the same command (shares, not absolutes).
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --recon-iters 0 > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err; echo "ncu rc=$?"
tail -n 3 gpurun_out/bench_c2.json gpurun_out/bench_ref.json; tail -n 5 gpurun_out/bench_c2.err
