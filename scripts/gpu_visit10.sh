#!/bin/bash
# Full GPU suite incl. the new edge/full-size/CLI tests, golden ext fixture, ncu --set full of the HBM-bound kernels.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s --durations=12 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -n "^\[\|passed\|failed\|rc=\|^E  \|Error\|^real" gpurun_out/pytest_gpu.log | tail -n 40
timeout 300 python tests/golden/make_golden.py ext 2>&1 | tail -n 1
# ncu --set full: forward chi2 kernel and image prep at C2 (one launch each), gridding merge at C5 x 0.05
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_degrid_chi2 -s 4 -c 1 -o gpurun_out/k_degrid_chi2_c2 \
  python bench.py --steps 1 --warmup 3 --recon-iters 0 --no-cpu-baseline > gpurun_out/ncu_degrid.log 2>&1; echo "ncu degrid rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_grid_accumulate -c 1 -o gpurun_out/k_grid_accumulate_c5_0p05 \
  python bench.py --config c5 --scale 0.05 --steps 1 --warmup 3 --recon-iters 0 --no-cpu-baseline > gpurun_out/ncu_gridacc.log 2>&1; echo "ncu gridacc rc=$?"
ls -la gpurun_out/*.ncu-rep
