#!/bin/bash
# Tile-sequential gridding: bit-exactness (vs the CPU reference order and vs the merge kernel), timing, ncu.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "gridd or weights or scenario" ) > gpurun_out/pytest_v26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v26.log
grep -n "passed\|failed\|rc=\|^E  \|Error" gpurun_out/pytest_v26.log | tail -n 8
GVM_PROFILE_HOST=1 timeout 1500 python bench.py --config c5 --scale 0.25 --steps 3 --warmup 3 --recon-iters 10 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c5.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["preprocessing"], d["recon"]["seconds"])
PY
grep "gvm_grid_block\|gvm_weights" gpurun_out/bench_c5.err | head -n 3
timeout 1500 python bench.py --config c4 --steps 3 --warmup 3 --recon-iters 10 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c4.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["preprocessing"], d["recon"]["seconds"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5.csv \
  python bench.py --config c5 --scale 0.05 --steps 1 --warmup 3 --recon-iters 0 --no-cpu-baseline > gpurun_out/ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
python scripts/launch_summary.py gpurun_out/launches_c5.csv > gpurun_out/launches_c5_summary.txt; grep "k_tile\|k_grid\|Onesweep\|Scan\|total" gpurun_out/launches_c5_summary.txt
