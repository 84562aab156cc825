#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_cli_gpu.py tests/test_host_gpu.py -m gpu -q -x ) > gpurun_out/pytest_v25.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v25.log
grep -n "passed\|failed\|rc=\|^E  \|Error" gpurun_out/pytest_v25.log | tail -n 10
for m in 1 0; do
  GVM_SINGLE_SYNC=$m GVM_PROFILE_HOST=1 timeout 600 python bench.py --config c1 --steps 3 --warmup 3 --recon-iters 50 --no-cpu-baseline > gpurun_out/bench_c1_sync$m.json 2> gpurun_out/bench_c1_sync$m.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_c1_sync$m.json").read().strip().splitlines()[-1])
print("GVM_SINGLE_SYNC=$m", d["ms_per_step"], d["recon"]["seconds"], d["recon"]["seconds_in_function_evals"], d["recon"]["function_evals"])
PY
done
grep -v "^$" gpurun_out/bench_c1_sync1.err | sed -n 2,8p
