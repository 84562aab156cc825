#!/bin/bash
# 2-GPU visit: sharded parity incl. error maps, the N=2 bench of both arms, the default bench with the cpu_gridding leg.
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -s ) > gpurun_out/pytest_mgpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mgpu.log
grep -n "^\[\|passed\|failed\|skipped\|rc=\|^E  \|Error" gpurun_out/pytest_mgpu.log | tail -n 12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; echo "n2 rc=$?"
tail -n 1 gpurun_out/bench_c2_n2.json | cut -c 1-700
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"
tail -n 1 gpurun_out/bench_ref_n2.json | cut -c 1-300
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]); print(d.get("cpu_gridding")); print(d["cpu_baseline"])
PY
tail -n 3 gpurun_out/bench_c2.err
