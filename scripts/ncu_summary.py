"""Compact text summary of an .ncu-rep (headline metrics + top stall sites) for profiles/.  Runs without a GPU.
usage: python scripts/ncu_summary.py gpurun_out/prof_chain.ncu-rep profiles/r1_chain_full.txt"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors_op_read.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum")


def run(*a):
    return subprocess.run(["ncu", "-i", rep, *a], capture_output=True, text=True).stdout


lines = []
rows = list(csv.reader(io.StringIO(run("--page", "raw", "--csv"))))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    lines.append(f"== launch {d.get('ID')}: {d.get('Kernel Name')}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
    for k in hdr:
        name = k.split(".", 2)[-1] if k.startswith(("SM_", "TPC.", "GPC.")) else k
        if any(name == x or k.endswith(x) for x in KEEP):
            lines.append(f"   {k} = {d[k]} {units[hdr.index(k)]}")
rows = list(csv.reader(io.StringIO(run("--page", "source", "--csv"))))
hdr, data, kern = None, [], 0
for r in rows:
    if r and r[0] == "Kernel Name":
        kern += 1
    elif r and r[0] == "Address":
        hdr = r
    elif kern == 1 and hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
tot = sum(int(d["# Samples"] or 0) for d in data)
lines.append(f"== warp-state samples, first captured launch: {tot} total; top sites (SASS, samples, executed, top stall reasons)")
for d in sorted(data, key=lambda d: -int(d["# Samples"] or 0))[:30]:
    st = {k: int(d[k]) for k in hdr if k.startswith("stall_") and "Not Issued" not in k and d[k] and int(d[k]) > 0}
    st = sorted(st.items(), key=lambda x: -x[1])[:3]
    lines.append(f"   {d['Address'][-6:]} {d['# Samples']:>6} {d['Instructions Executed']:>9}  {d['Source'][:70]:70s} {st}")
open(out, "w").write("\n".join(lines) + "\n")
print(f"wrote {out} ({len(lines)} lines)")
