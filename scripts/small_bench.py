"""Small-batch chain timing: weights-stationary cluster kernel vs the tcgen05 kernel (hopper / walker2d, a few envs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dppo_b200.workloads import get_workload
from tests.helpers import build_model, our_classes

for name, Es in (("hopper", (1, 8, 40, 48)), ("walker2d", (40,))):
    w = get_workload(name)
    model = build_model(w, "cuda:0", our_classes())
    eng = model.engine()
    for E in Es:
        st = torch.rand(E, 1, w["obs_dim"], device="cuda") * 2 - 1
        for label, shape in (("small", (0, -1)), ("tcgen05 NE=16 C=4", (16, 4))):
            eng.set_launch_shape(*shape)
            try:
                for _ in range(5):
                    eng.sample(st)
            except RuntimeError as ex:
                print(name, E, label, "unavailable:", ex)
                continue
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                eng.sample(st)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 20
            print(f"{name} E={E} {label}: {ms * 1e3:.1f} us/chain  {E * w['act_steps'] / ms * 1e3 / 1e3:.0f} k env-steps/s")
