#!/bin/bash
# one GPU call: parity tests, role counters, micro-benchmarks, bench, ncu launch list + full capture of the chain kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for ne in 64 32; do DPPO_B200_TILE_ENVS=$ne timeout 120 python scripts/chain_prof.py walker2d 4096 split3; done > gpurun_out/chain_prof.log 2>&1
timeout 300 python scripts/microbench.py > gpurun_out/microbench.log 2>&1; echo "microbench rc=$?"
BENCH="python bench.py --steps 5 --warmup 3"
timeout 600 $BENCH > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err
rc=$?; echo "bench rc=$rc"; cat gpurun_out/bench_plain.json
if [ $rc -eq 0 ]; then
  SHORT="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
  timeout 300 $SHORT > gpurun_out/short_plain.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_list.log 2>&1
  echo "ncu list rc=$?"
  SHORT2="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-update"
  timeout 300 $SHORT2 > gpurun_out/short2_plain.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_mlp -s 3 -c 2 -f -o gpurun_out/prof_chain $SHORT2 > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi
