#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c4.csv \
  python bench.py --config c4 --scale 0.2 --steps 1 --warmup 3 --recon-iters 0 --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
python scripts/launch_summary.py gpurun_out/launches_c4.csv > gpurun_out/launches_c4_summary.txt; grep "k_tile\|k_grid\|Onesweep\|Scan\|total" gpurun_out/launches_c4_summary.txt
GVM_PROFILE_HOST=1 timeout 900 python bench.py --config c4 --steps 1 --warmup 3 --recon-iters 0 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; grep "gvm_grid\|gvm_weights\|gvm_add_channel" gpurun_out/bench_c4.err | head -5
