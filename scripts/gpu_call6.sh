#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload square_unet > gpurun_out/bench_r1c_unet.json 2> gpurun_out/bench_r1c_unet.err; echo "bench unet rc=$?"
cat gpurun_out/bench_r1c_unet.json; tail -3 gpurun_out/bench_r1c_unet.err
timeout 600 python bench.py --workload square_unet --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1c_unet_ref.json 2> gpurun_out/bench_r1c_unet_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_r1c_unet_ref.json
