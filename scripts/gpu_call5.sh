#!/bin/bash
# round-1b evidence: GPU tests, plain bench, ncu launch list, one full capture of the chain kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; echo "bench rc=$?"
cat gpurun_out/bench_r1b.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1b_ref.json 2> gpurun_out/bench_r1b_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_r1b_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-update > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_mlp -c 1 -s 3 -o gpurun_out/prof_chain_r1b -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-update > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
