"""Where the host-side time of one e2e rollout decision goes at small E (hopper, 40 envs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dppo_b200.workloads import get_workload
from tests.helpers import build_model, our_classes

w = get_workload("hopper"); E = 40
model = build_model(w, "cuda:0", our_classes())
dev = torch.device("cuda:0")
obs_h = torch.rand(E, 1, w["obs_dim"]).pin_memory()
ft = w["ft_denoising_steps"]
h_traj = torch.empty((E, w["horizon_steps"], w["action_dim"])).pin_memory()
h_chain = torch.empty((E, ft + 1, w["horizon_steps"], w["action_dim"])).pin_memory()
def step():
    obs = obs_h.to(dev, non_blocking=True)
    out = model(cond={"state": obs})
    h_traj.copy_(out.trajectories, non_blocking=True)
    h_chain.copy_(out.chains, non_blocking=True)
    torch.cuda.current_stream().synchronize()
for _ in range(20): step()
N = 200
t0 = time.perf_counter()
for _ in range(N): step()
print(f"e2e step: {(time.perf_counter() - t0) / N * 1e6:.1f} us")
# pieces (host time only, device idle between)
def timeit(fn, n=N):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    t = (time.perf_counter() - t0) / n * 1e6; torch.cuda.synchronize(); return t
obs = obs_h.to(dev)
print(f"  obs.to(dev) host cost        {timeit(lambda: obs_h.to(dev, non_blocking=True)):.1f} us")
print(f"  model.engine() (weight sync) {timeit(lambda: model.engine()):.1f} us")
eng = model.engine()
print(f"  eng.sample host cost         {timeit(lambda: eng.sample(obs)):.1f} us (includes kernel time when the queue backs up)")
out = model(cond={"state": obs})
print(f"  2 x D2H copy_ host cost      {timeit(lambda: (h_traj.copy_(out.trajectories, non_blocking=True), h_chain.copy_(out.chains, non_blocking=True))):.1f} us")
def full_sync():
    model(cond={"state": obs}); torch.cuda.current_stream().synchronize()
print(f"  model() + sync               {timeit(full_sync):.1f} us")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
