#!/bin/bash
mkdir -p gpurun_out
python scripts/diag/fn_eval_timing.py c1 2>&1 | grep -v "^$" | tail -n 8 | tee gpurun_out/fn_eval_timing_c1.txt
python scripts/diag/fn_eval_timing.py c2 2>&1 | tail -n 7 | tee gpurun_out/fn_eval_timing_c2.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_c5.csv \
    python bench.py --config c5 --scale 0.05 --steps 1 --warmup 3 --no-cpu-baseline --recon-iters 0 > gpurun_out/bench_c5_ncu.json 2>&1
python scripts/launch_summary.py gpurun_out/launches_c5.csv "C5 (8192^2 grid, 10M raw visibilities -> 2.4M gridded samples), bench.py --config c5 --scale 0.05" | tail -n 30
