#!/bin/bash
# Cached attenuation planes: whole GPU suite (bit-exactness vs the reference), then C5 (1/4), C4, C1, C2 benches.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -n "passed\|failed\|rc=\|^E  \|Error\|^real" gpurun_out/pytest_gpu.log | tail -n 8
for cfg in "c5 0.25 10" "c4 1.0 10" "c1 1.0 50"; do
  set -- $cfg
  timeout 1500 python bench.py --config $1 --scale $2 --steps 5 --warmup 3 --recon-iters $3 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "$1 rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$1.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["recon"]["seconds"], d["recon"]["seconds_in_function_evals"], d["recon"]["function_evals"])
PY
done
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["recon"]["seconds"], d["gpu_launches"])
PY
