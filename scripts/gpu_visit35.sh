#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -s ) > gpurun_out/pytest_mgpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mgpu.log
grep -n "^\[\|passed\|failed\|skipped\|rc=\|^E  \|Error" gpurun_out/pytest_mgpu.log | tail -n 8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; echo "n2 rc=$?"
grep -h '^{' gpurun_out/bench_c2_n2.json | cut -c 1-260
