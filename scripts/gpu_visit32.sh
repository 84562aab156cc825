#!/bin/bash
mkdir -p gpurun_out
for cfg in "c4 1.0" "c5 0.25"; do
  set -- $cfg
  GVM_GRID_TIMING=1 GVM_PROFILE_HOST=1 timeout 900 python bench.py --config $1 --scale $2 --steps 1 --warmup 3 --recon-iters 0 --no-cpu-baseline > gpurun_out/bench_t_$1.json 2> gpurun_out/bench_t_$1.err; echo "$1 rc=$?"
  grep "gvm timing\|gvm_grid_block\|gvm_weights" gpurun_out/bench_t_$1.err | head -n 12
done
