#!/bin/bash
# Mosaic test + C5 at FULL scale (200 M raw visibilities) if the box has the host memory for the generator.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_edges_gpu.py -m gpu -q -s -k "mosaic" ) > gpurun_out/pytest_v15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v15.log
grep -n "passed\|failed\|rc=\|^E  \|Error" gpurun_out/pytest_v15.log | tail -n 12
grep -E "MemTotal|MemAvailable" /proc/meminfo; nproc
avail=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
if [ "$avail" -ge 96 ]; then
  ( time timeout 1200 python bench.py --config c5 --scale 1.0 --steps 3 --warmup 3 --recon-iters 10 --no-cpu-baseline ) > gpurun_out/bench_c5_full.json 2> gpurun_out/bench_c5_full.err; echo "c5 full rc=$?"
  python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c5_full.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["preprocessing"], d["recon"])
PY
  tail -n 4 gpurun_out/bench_c5_full.err
else
  echo "only $avail GiB of host memory available: C5 full scale skipped"
fi
