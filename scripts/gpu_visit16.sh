#!/bin/bash
# 8-GPU visit: the default bench (C2, visibility chunks) and C3 (64 channels, i % 8) at N = 8.
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( time timeout 600 $TR --master-port 29701 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_c2_n8.json 2> gpurun_out/bench_c2_n8.err; echo "c2 n8 rc=$?"
grep -h '^{' gpurun_out/bench_c2_n8.json | cut -c 1-400; tail -n 3 gpurun_out/bench_c2_n8.err
( time timeout 900 $TR --master-port 29702 bench.py --config c3 --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline --recon-iters 3 ) > gpurun_out/bench_c3_n8.json 2> gpurun_out/bench_c3_n8.err; echo "c3 n8 rc=$?"
grep -h '^{' gpurun_out/bench_c3_n8.json | cut -c 1-400; tail -n 3 gpurun_out/bench_c3_n8.err
