#!/bin/bash
# evidence recipe of a round: default bench (both arms), ncu launch list of the same command, one full capture per chain kernel and of the update kernels.  usage: gpurun -- bash scripts/gpu_evidence.sh (edit the r1l tag)
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_r1l.json 2> gpurun_out/bench_r1l.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1l_ref.json 2>/dev/null; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1l.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_mlp -c 1 -s 3 -o gpurun_out/prof_chain_mlp_r1l -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-update > gpurun_out/ncu_full_mlp.log 2>&1; echo "ncu mlp rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_unet -c 1 -s 3 -o gpurun_out/prof_chain_unet_r1l -f python bench.py --workload square_unet --steps 2 --warmup 3 --no-cpu-baseline --no-update > gpurun_out/ncu_full_unet.log 2>&1; echo "ncu unet rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_small -c 1 -s 3 -o gpurun_out/prof_chain_small_r1l -f python bench.py --workload hopper --steps 2 --warmup 3 --no-cpu-baseline --no-update > gpurun_out/ncu_full_small.log 2>&1; echo "ncu small rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"ppo_loss|adamw|gae|adv_stats" -c 6 -o gpurun_out/prof_update_r1l -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_update.log 2>&1; echo "ncu update rc=$?"
for w in transport furniture square_unet hopper; do timeout 400 python bench.py --workload $w > gpurun_out/bench_r1l_$w.json 2>/dev/null; echo "bench $w rc=$?"; done
ls -la gpurun_out/*r1l*.ncu-rep
