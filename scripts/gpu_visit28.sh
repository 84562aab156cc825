#!/bin/bash
# Final validation: full GPU suite, ncu --set full of the tile gridding kernel.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -n "passed\|failed\|rc=\|^E  \|Error\|^real" gpurun_out/pytest_gpu.log | tail -n 6
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_grid_tiles -c 1 -o gpurun_out/k_grid_tiles_c5_0p05 \
  python bench.py --config c5 --scale 0.05 --steps 1 --warmup 3 --recon-iters 0 --no-cpu-baseline > gpurun_out/ncu_gridtiles.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/k_grid_tiles_c5_0p05.ncu-rep
