"""Phase cycle counters of the small-batch chain kernel (needs a DPPO_B200_CHAIN_PROF=1 build).  python scripts/small_prof.py [workload] [envs]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from dppo_b200 import _lib
from dppo_b200.workloads import get_workload
from tests.helpers import build_model, our_classes

name = sys.argv[1] if len(sys.argv) > 1 else "hopper"
w = get_workload(name)
E = int(sys.argv[2]) if len(sys.argv) > 2 else w["n_envs"]
model = build_model(w, "cuda:0", our_classes())
eng = model.engine()
eng.set_launch_shape(0, -1)
prof = torch.zeros(256 * 16, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.dppo_debug_set_prof.argtypes = [C.c_void_p, C.c_void_p]
lib.dppo_debug_set_prof(eng.ctx, C.c_void_p(prof.data_ptr()))
state = torch.rand(E, 1, w["obs_dim"], device="cuda") * 2 - 1
for _ in range(3):
    eng.sample(state)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
eng.sample(state)
b.record()
torch.cuda.synchronize()
p = prof.view(256, 16).cpu().double()
used = p[p[:, 4] > 0]
S = w["denoising_steps"]
print(f"{name} E={E}: kernel {a.elapsed_time(b):.4f} ms, {len(used)} CTAs, per step (thread 0, mean over CTAs), cycles:")
for i, n in enumerate(["wait for a buffer (4 per step)", "dot products (2 hidden layers)", "reduce + publish", "output layer + posterior", "total"]):
    print(f"  {n:34s} {used[:, i].mean() / S:9.0f}   max {used[:, i].max() / S:9.0f}")
