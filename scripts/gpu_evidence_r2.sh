#!/bin/bash
# Round-2 evidence: launch list of the default bench, one `--set full` capture per chain kernel and of the update GEMMs of
# one walker2d / furniture minibatch.  The captures are summarised ON the box (scripts/ncu_summary.py, raw csv) and only
# the chain_mlp report travels back: gpurun merges at most 64 MiB.   usage: gpurun --timeout 1800 -- bash scripts/gpu_evidence_r2.sh
mkdir -p gpurun_out /tmp/rep
T=r2k
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong"
timeout 300 $B > gpurun_out/${T}_plain_bench.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv $B > gpurun_out/${T}_ncu_list.log 2>&1; echo "launch list rc=$?"
C="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --no-update"
timeout 300 $C > /dev/null 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_mlp -c 1 -s 3 -o gpurun_out/${T}_prof_chain_mlp -f $C > gpurun_out/${T}_ncu_mlp.log 2>&1; echo "ncu mlp rc=$?"
python scripts/ncu_summary.py gpurun_out/${T}_prof_chain_mlp.ncu-rep gpurun_out/${T}_chain_mlp_ncu_full.txt
timeout 300 $C --workload hopper > /dev/null 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_small -c 1 -s 3 -o /tmp/rep/small -f $C --workload hopper > gpurun_out/${T}_ncu_small.log 2>&1; echo "ncu small rc=$?"
python scripts/ncu_summary.py /tmp/rep/small.ncu-rep gpurun_out/${T}_chain_small_ncu_full.txt
timeout 300 $C --workload square_unet > /dev/null 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_unet -c 1 -s 3 -o /tmp/rep/unet -f $C --workload square_unet > gpurun_out/${T}_ncu_unet.log 2>&1; echo "ncu unet rc=$?"
python scripts/ncu_summary.py /tmp/rep/unet.ncu-rep gpurun_out/${T}_chain_unet_ncu_full.txt
for w in walker2d furniture; do
  U="python scripts/update_perf.py --workload $w --reps 1 --no-graph --n-steps 10 --profile-range"
  timeout 300 $U > gpurun_out/${T}_plain_update_$w.log 2>&1 && timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"ugemm|ln_fwd|ln_bwd" -c 40 -o /tmp/rep/update_$w -f $U > gpurun_out/${T}_ncu_update_$w.log 2>&1; echo "ncu update $w rc=$?"
  ncu -i /tmp/rep/update_$w.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size > gpurun_out/${T}_update_kernels_$w.csv 2>&1
done
du -sh gpurun_out; ls -la gpurun_out/${T}_*
