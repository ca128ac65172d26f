"""GPU bring-up of the Unet1D chain kernel: chains / log-probs of cfg5 against the golden vectors, then timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dppo_b200.workloads import get_workload
from tests.helpers import GOLDEN_CASES, build_model, load_golden, make_inputs, our_classes

case = "square_unet"
spec = GOLDEN_CASES[case]
w = get_workload(spec["workload"])
model = build_model(w, "cuda:0", our_classes())
gold = load_golden(case)
inp = make_inputs(w, spec["n_envs"], spec["mb_rows"])
state, noise = inp["state"].cuda(), inp["noise"].cuda()


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(1.0, np.abs(b))


for ne, cc in ((16, 1), (16, 2), (32, 1), (32, 2)):
    model.engine().set_launch_shape(ne, cc)
    try:
        out = model(cond={"state": state}, deterministic=False, return_chain=True, noise=noise)
    except RuntimeError as ex:
        print(f"NE={ne} C={cc}: {ex}")
        continue
    torch.cuda.synchronize()
    e = rel(out.chains.cpu().numpy(), gold["chains"])
    print(f"NE={ne} C={cc} chains max rel err {e.max():.3e}  per-slot max {e.reshape(e.shape[0], e.shape[1], -1).max(axis=(0, 2))}")
    with torch.no_grad():
        lp = model.get_logprobs({"state": state}, torch.from_numpy(gold["chains"]).cuda())
    e = rel(lp.cpu().numpy(), gold["logprobs"])
    print(f"NE={ne} C={cc} logprobs max rel err {e.max():.3e}")
    out = model(cond={"state": state}, deterministic=True, return_chain=True, noise=noise)
    e = rel(out.chains.cpu().numpy(), gold["chains_det"])
    print(f"NE={ne} C={cc} deterministic chains max rel err {e.max():.3e}")

model.engine().set_launch_shape(0, 0)
for E in (1024, 2048, 4096):
    st = torch.rand(E, 1, w["obs_dim"], device="cuda") * 2 - 1
    for ne, cc in ((16, 1), (16, 2), (32, 1), (32, 2), (0, 0)):
        model.engine().set_launch_shape(ne, cc)
        try:
            model(cond={"state": st})
        except RuntimeError:
            continue
        for _ in range(3):
            model(cond={"state": st})
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            model(cond={"state": st})
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 10
        print(f"E={E} NE={ne} C={cc}: {ms:.3f} ms/chain  {E * w['act_steps'] / ms * 1e3 / 1e6:.2f} M env-steps/s")
