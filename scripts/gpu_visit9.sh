#!/bin/bash
# Error maps + conv degridding parity, golden ext fixture, gridding after the host/kernel changes.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -s -k "error or degrid or gridd or weights" ) > gpurun_out/pytest_v9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v9.log
grep -n "^\[\|error maps\|passed\|failed\|rc=\|Error\|assert" gpurun_out/pytest_v9.log | tail -n 30
timeout 300 python tests/golden/make_golden.py ext 2>&1 | tail -n 2
GVM_PROFILE_HOST=1 timeout 1500 python bench.py --config c5 --scale 0.25 --steps 3 --warmup 3 --recon-iters 10 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c5.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["preprocessing"], d["recon"]["seconds"])
PY
grep -v "^$" gpurun_out/bench_c5.err | head -n 40
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_c5.csv \
  python bench.py --config c5 --scale 0.05 --steps 2 --warmup 3 --recon-iters 0 --no-cpu-baseline > gpurun_out/ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
python scripts/launch_summary.py gpurun_out/launches_c5.csv > gpurun_out/launches_c5_summary.txt; grep "k_grid\|Onesweep\|total" gpurun_out/launches_c5_summary.txt
