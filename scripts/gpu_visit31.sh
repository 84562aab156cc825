#!/bin/bash
# Pipelined host copies: whole GPU suite, then preprocessing timings (C5 1/4 and full, C4, C2 setup).
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -n "passed\|failed\|rc=\|^E  \|Error\|^real" gpurun_out/pytest_gpu.log | tail -n 6
for cfg in "c5 0.25" "c5 1.0" "c4 1.0"; do
  set -- $cfg
  GVM_PROFILE_HOST=1 timeout 1500 python bench.py --config $1 --scale $2 --steps 3 --warmup 3 --recon-iters 10 --no-cpu-baseline > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err; echo "$1 $2 rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$1_$2.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["preprocessing"], d["recon"]["seconds"], d["recon"]["setup_seconds"])
PY
  grep "gvm_grid_block\|gvm_weights\|gvm_add_channel" gpurun_out/bench_$1_$2.err | head -n 3
done
