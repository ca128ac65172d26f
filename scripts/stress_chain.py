"""Back-to-back launches of the chain kernel to expose rare protocol races.  usage: stress_chain.py workload E NE C launches"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dppo_b200.workloads import get_workload
from tests.helpers import build_model, our_classes

name, E, ne, c, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
w = get_workload(name)
model = build_model(w, "cuda:0", our_classes())
eng = model.engine()
eng.set_launch_shape(ne, c)
st = torch.rand(E, 1, w["obs_dim"], device="cuda") * 2 - 1
t0 = time.perf_counter()
try:
    for i in range(n):
        eng.sample(st, seed=1, offset=i + 1)
        if i % 64 == 63:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    print(f"{name} E={E} NE={ne} C={c}: {n} launches OK in {time.perf_counter() - t0:.1f} s", flush=True)
except Exception as ex:
    print(f"{name} E={E} NE={ne} C={c}: FAILED after <= {i} launches: {str(ex)[:120]}", flush=True)
