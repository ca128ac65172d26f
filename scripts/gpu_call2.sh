#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
for wl in "walker2d 4096" "hopper 40" "transport 50" "furniture 1000" "furniture 125" "transport_k20 50"; do
  timeout 300 python scripts/shape_sweep.py $wl
done > gpurun_out/shape_sweep.log 2>&1
cat gpurun_out/shape_sweep.log
