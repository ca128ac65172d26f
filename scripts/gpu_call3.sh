#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
for sh in "64 2" "32 1" "64 1"; do timeout 120 python scripts/chain_prof.py walker2d 4096 split3 $sh; done 2>&1 | grep -E "kernel|mma_wait|mma_total|epi_hand|warp2|warp3"
for wl in "walker2d 4096" "hopper 40" "furniture 1000"; do timeout 300 python scripts/shape_sweep.py $wl; done 2>&1 | tee gpurun_out/shape_sweep.log
