#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "gridd or scenario" ) > gpurun_out/pytest_v29.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v29.log
grep -n "passed\|failed\|rc=\|^E  \|Error" gpurun_out/pytest_v29.log | tail -n 6
for cfg in "c5 0.05" "c4 0.2"; do
  set -- $cfg
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$1.csv \
    python bench.py --config $1 --scale $2 --steps 1 --warmup 3 --recon-iters 0 --no-cpu-baseline > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
  python scripts/launch_summary.py gpurun_out/launches_$1.csv > gpurun_out/launches_$1_summary.txt; grep "k_tile\|k_grid_tiles\|Onesweep\|total" gpurun_out/launches_$1_summary.txt
done
