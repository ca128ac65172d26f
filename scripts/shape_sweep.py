"""Chain-kernel time for every launch shape (tile envs x cluster size).  python scripts/shape_sweep.py workload [envs]"""
import sys

import torch

sys.path.insert(0, ".")
from dppo_b200.workloads import get_workload
from tests.helpers import build_model, our_classes

name = sys.argv[1] if len(sys.argv) > 1 else "walker2d"
w = get_workload(name)
E = int(sys.argv[2]) if len(sys.argv) > 2 else w["n_envs"]
model = build_model(w, "cuda:0", our_classes())
eng = model.engine()
state = torch.rand(E, 1, w["obs_dim"], device="cuda") * 2 - 1
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
MT = w["actor"]["mlp_dims"][0] // 128
for ne in (16, 32, 64):
    if ne == 64 and MT > 4:
        continue
    for c in (1, 2, 4, 8):
        if MT % c:
            continue
        eng.set_launch_shape(ne, c)
        try:
            for _ in range(2):
                eng.sample(state)
            torch.cuda.synchronize()
            ts = []
            for i in range(5):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                eng.sample(state, offset=i)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            print(f"{name} E={E} NE={ne} C={c}: {min(ts):.3f} ms (median {sorted(ts)[2]:.3f})  grid {((E + ne - 1) // ne) * c}", flush=True)
        except RuntimeError as e:
            print(f"{name} E={E} NE={ne} C={c}: FAILED {e}", flush=True)
            break
eng.set_launch_shape(0, 0)
ts = []
for i in range(5):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.sample(state, offset=i)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"{name} E={E} auto shape: {min(ts):.3f} ms")
