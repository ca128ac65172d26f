#!/bin/bash
mkdir -p gpurun_out
for wl in hopper transport furniture; do
  timeout 600 python bench.py --workload $wl --no-update > gpurun_out/bench_r1d_$wl.json 2> gpurun_out/bench_r1d_$wl.err; echo "$wl rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_r1d_$wl.json"))
print("$wl", "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "e2e_ms", round(d["e2e"]["ms_per_step"],4), "cpu", round(d["cpu_baseline"]["value"]), "ratio_e2e", round(d["e2e"]["value"]/d["cpu_baseline"]["value"],1), "frac", round(d["roofline"]["frac"],4))
PY
done
