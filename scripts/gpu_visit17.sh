#!/bin/bash
# Half-plane forward model + C2R gridded gradient: tests, then C5 (1/4) and C4 benches.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -s -k "half_plane or gridd or error_maps or scenario" ) > gpurun_out/pytest_v17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v17.log
grep -n "gridded gradient\|passed\|failed\|rc=\|^E  \|Error" gpurun_out/pytest_v17.log | tail -n 14
for cfg in "c5 0.25" "c4 1.0"; do
  set -- $cfg
  timeout 1500 python bench.py --config $1 --scale $2 --steps 5 --warmup 3 --recon-iters 10 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "$1 rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$1.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["recon"]["seconds"], d["recon"]["seconds_in_function_evals"], d["recon"]["function_evals"])
PY
  tail -n 2 gpurun_out/bench_$1.err
done
