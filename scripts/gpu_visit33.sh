#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "gridd or weights or scenario" ) > gpurun_out/pytest_v33.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v33.log
grep -n "passed\|failed\|rc=\|^E  \|Error" gpurun_out/pytest_v33.log | tail -n 6
for cfg in "c5 0.25" "c5 1.0"; do
  set -- $cfg
  GVM_PROFILE_HOST=1 timeout 1500 python bench.py --config $1 --scale $2 --steps 3 --warmup 3 --recon-iters 10 --no-cpu-baseline > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err; echo "$1 $2 rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$1_$2.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["preprocessing"], d["recon"]["seconds"])
PY
  grep "gvm_grid_block\|gvm_weights" gpurun_out/bench_$1_$2.err | head -n 2
done
