#!/bin/bash
# Round-2 evidence: GPU test suite, launch list of the default bench, one `--set full` capture per dominant kernel
# (chain kernels + the update GEMMs of one walker2d / furniture minibatch).  usage: gpurun --timeout 2400 -- bash scripts/gpu_evidence_r2.sh
mkdir -p gpurun_out
T=r2k
timeout 700 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${T}_pytest_gpu.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong"
timeout 300 $B > gpurun_out/${T}_plain_bench.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv $B > gpurun_out/${T}_ncu_list.log 2>&1; echo "launch list rc=$?"
C="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --no-update"
timeout 300 $C > /dev/null 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_mlp -c 1 -s 3 -o gpurun_out/${T}_prof_chain_mlp -f $C > gpurun_out/${T}_ncu_mlp.log 2>&1; echo "ncu mlp rc=$?"
timeout 300 $C --workload hopper > /dev/null 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_small -c 1 -s 3 -o gpurun_out/${T}_prof_chain_small -f $C --workload hopper > gpurun_out/${T}_ncu_small.log 2>&1; echo "ncu small rc=$?"
timeout 300 $C --workload square_unet > /dev/null 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_unet -c 1 -s 3 -o gpurun_out/${T}_prof_chain_unet -f $C --workload square_unet > gpurun_out/${T}_ncu_unet.log 2>&1; echo "ncu unet rc=$?"
for w in walker2d furniture; do
  U="python scripts/update_perf.py --workload $w --reps 1 --no-graph --n-steps 10 --profile-range"
  timeout 300 $U > gpurun_out/${T}_plain_update_$w.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ugemm|ln_fwd|ln_bwd|pack_rows|ppo_loss" -c 80 -o gpurun_out/${T}_prof_update_$w -f $U > gpurun_out/${T}_ncu_update_$w.log 2>&1; echo "ncu update $w rc=$?"
done
ls -la gpurun_out/${T}_*.ncu-rep
