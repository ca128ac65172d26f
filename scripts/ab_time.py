"""Kernel-only timing of one workload's chain launch (A/B aid): python scripts/ab_time.py [workload] [iters]"""
import sys

import torch

sys.path.insert(0, ".")
from dppo_b200.workloads import get_workload
from tests.helpers import build_model, our_classes

name = sys.argv[1] if len(sys.argv) > 1 else "walker2d"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
w = get_workload(name)
E = w["n_envs"]
model = build_model(w, "cuda:0", our_classes())
eng = model.engine()
state = torch.rand(E, 1, w["obs_dim"], device="cuda") * 2 - 1
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(10):
    eng.sample(state)
torch.cuda.synchronize()
ts = []
for _ in range(iters):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.sample(state)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
print(f"{name} E={E}: median {ts[len(ts) // 2]:.4f} ms  min {ts[0]:.4f}  p90 {ts[int(len(ts) * 0.9)]:.4f}")
