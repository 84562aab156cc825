#!/bin/bash
# ncu --set full capture of the dominant kernel on the C2 image shape (1M visibilities, unmasked)
mkdir -p gpurun_out
PROBE_BMAX=14000 PROBE_BMIN=150 PROBE_CHUNKS=2048 PROBE_NPIX=50 timeout 600 ncu --set full --clock-control none --import-source on \
   -k regex:k_grad_umma -s 1 -c 1 -o gpurun_out/umma_pair python scripts/umma_probe.py > gpurun_out/umma_pair_ncu.log 2>&1
tail -n 3 gpurun_out/umma_pair_ncu.log
