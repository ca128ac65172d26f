#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log | head -5
for wl in "furniture 1000" "furniture 125" "furniture_ddpm100 125"; do timeout 300 python scripts/shape_sweep.py $wl; done 2>&1 | tee gpurun_out/shape_sweep2.log
