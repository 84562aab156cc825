#!/bin/bash
# Filter test; where a function evaluation spends its time (C1 host profile, C5 launch list after the half-plane change).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cli_gpu.py -m gpu -q 2>&1 | tail -n 3
GVM_PROFILE_HOST=1 timeout 600 python bench.py --config c1 --steps 3 --warmup 3 --recon-iters 50 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1_hostprofile.txt; echo "c1 rc=$?"
grep -v "^$" gpurun_out/bench_c1_hostprofile.txt | head -n 14
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_c5.csv \
  python bench.py --config c5 --scale 0.05 --steps 2 --warmup 3 --recon-iters 0 --no-cpu-baseline > gpurun_out/ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
python scripts/launch_summary.py gpurun_out/launches_c5.csv > gpurun_out/launches_c5_summary.txt; grep -v "k_weight\|cub::\|k_cell\|k_clear\|k_grid_centres\|k_grid_flags\|k_grid_compact\|k_prep\|k_max\|k_min\|k_noise" gpurun_out/launches_c5_summary.txt
