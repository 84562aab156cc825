#!/bin/bash
# 4-GPU point of the scaling table (C2 default bench), both arms.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29801 bench.py --impl reference --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_ref_n4.json 2> gpurun_out/bench_ref_n4.err; echo "ref n4 rc=$?"
timeout 600 $TR --master-port 29802 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_n4.json 2> gpurun_out/bench_c2_n4.err; echo "c2 n4 rc=$?"
grep -h '^{' gpurun_out/bench_ref_n4.json | cut -c 1-200
grep -h '^{' gpurun_out/bench_c2_n4.json | cut -c 1-330; tail -n 2 gpurun_out/bench_c2_n4.err
