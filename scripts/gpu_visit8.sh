#!/bin/bash
# Error-map parity + the other BASELINE configs (C1, C3, C4, C5) + launch lists of the HBM-bound paths.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -s -k "error or scenario" ) > gpurun_out/pytest_errmaps.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_errmaps.log
grep -n "^\[\|error maps\|passed\|failed\|rc=\|Error\|assert" gpurun_out/pytest_errmaps.log | tail -n 30
for cfg in "c1 1.0 50" "c3 1.0 5" "c4 1.0 10" "c5 0.25 10"; do
  set -- $cfg
  ( time timeout 1500 python bench.py --config $1 --scale $2 --steps 3 --warmup 3 --recon-iters $3 --no-cpu-baseline ) > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "$1 rc=$?"
  tail -n 1 gpurun_out/bench_$1.json | cut -c 1-1800; tail -n 4 gpurun_out/bench_$1.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_c5.csv \
  python bench.py --config c5 --scale 0.05 --steps 2 --warmup 3 --recon-iters 0 --no-cpu-baseline > gpurun_out/ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_c1.csv \
  python bench.py --config c1 --steps 2 --warmup 3 --recon-iters 0 --no-cpu-baseline > gpurun_out/ncu_c1.log 2>&1; echo "ncu c1 rc=$?"
python scripts/launch_summary.py gpurun_out/launches_c5.csv | head -n 30
python scripts/launch_summary.py gpurun_out/launches_c1.csv | head -n 20
