#!/bin/bash
# Round-end style validation: full GPU suite, smoke, both bench arms, launch list of the default bench command.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -n "passed\|failed\|rc=\|^E  \|Error\|^real" gpurun_out/pytest_gpu.log | tail -n 6
( time timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log | head -n 1
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
r = json.loads(open("gpurun_out/bench_ref.json").read().strip().splitlines()[-1])
d = json.loads(open("gpurun_out/bench_c2.json").read().strip().splitlines()[-1])
print("ref", r["value"], r["ms_per_step"]); print("ours", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["recon"]["seconds"], d["gpu_launches"], d["clocks"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_c2.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --recon-iters 0 > gpurun_out/ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
python scripts/launch_summary.py gpurun_out/launches_c2.csv "ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 3 --no-cpu-baseline --recon-iters 0 (C2, round 1d)" > gpurun_out/launches_c2_summary.txt; tail -n 12 gpurun_out/launches_c2_summary.txt
