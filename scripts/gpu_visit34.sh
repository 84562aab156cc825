#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -n "passed\|failed\|rc=\|^E  \|Error\|^real" gpurun_out/pytest_gpu.log | tail -n 6
for cfg in "c5 1.0" "c4 1.0"; do
  set -- $cfg
  GVM_PROFILE_HOST=1 timeout 1500 python bench.py --config $1 --scale $2 --steps 5 --warmup 3 --recon-iters 10 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "$1 $2 rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$1.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["preprocessing"], d["recon"]["seconds"])
PY
done
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["recon"]["seconds"], d["cpu_gridding"])
PY
