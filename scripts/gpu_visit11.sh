#!/bin/bash
# Gridding merge with shared-memory state: bit-exact tests, C5 x 0.25 bench + host profile, ncu full of the new kernel, C4 bench.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -s -k "gridd or weights or scenario" ) > gpurun_out/pytest_v11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v11.log
grep -n "passed\|failed\|rc=\|^E  \|Error" gpurun_out/pytest_v11.log | tail -n 12
GVM_PROFILE_HOST=1 timeout 1500 python bench.py --config c5 --scale 0.25 --steps 3 --warmup 3 --recon-iters 10 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c5.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["preprocessing"], d["recon"]["seconds"])
PY
grep -v "^$" gpurun_out/bench_c5.err | head -n 8
timeout 1500 python bench.py --config c4 --steps 3 --warmup 3 --recon-iters 10 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c4.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["preprocessing"], d["recon"]["seconds"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_grid_accumulate -c 1 -o gpurun_out/k_grid_accumulate_c5_0p05 \
  python bench.py --config c5 --scale 0.05 --steps 1 --warmup 3 --recon-iters 0 --no-cpu-baseline > gpurun_out/ncu_gridacc.log 2>&1; echo "ncu gridacc rc=$?"
