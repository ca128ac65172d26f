"""Launch shapes the chain kernel's cost model picks + the co-resident cluster counts it got from the occupancy calculator."""
import ctypes as C
import sys
import time

import torch

sys.path.insert(0, ".")
from dppo_b200.workloads import get_workload
from tests.helpers import build_model, our_classes

for name, sizes in (("walker2d", [4096, 2048, 512]), ("furniture", [1000, 500, 250, 125]), ("transport", [50, 25, 7]), ("hopper", [40])):
    w = get_workload(name)
    m = build_model(w, "cuda:0", our_classes())
    eng = m.engine()
    lib = eng.lib
    lib.dppo_debug_get_shape.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    tab = (C.c_int * 12)()
    for E in sizes:
        ne, c = C.c_int(), C.c_int()
        lib.dppo_debug_get_shape(eng.ctx, E, C.byref(ne), C.byref(c), tab)
        state = torch.rand(E, 1, w["obs_dim"], device="cuda") * 2 - 1
        for _ in range(3):
            eng.sample(state, seed=1, offset=1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            eng.sample(state, seed=1, offset=1)
        b.record()
        torch.cuda.synchronize()
        print(f"{name} E={E}: NE={ne.value} C={c.value}  {a.elapsed_time(b) / 20:.3f} ms   clusters[NE16/32/64][C1/2/4/8]={list(tab)}", flush=True)
