"""A/B of the host-buffer call on the Hopper case (40 envs, K = 20): interleaved blocks of model(cond={'state': host})
in one process; run once per setting of DPPO_B200_TICKET (the switch is read once per process)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dppo_b200.workloads import get_workload  # noqa: E402
from tests.helpers import build_model, our_classes  # noqa: E402

w = get_workload("hopper")
E = w["n_envs"]
model = build_model(w, "cuda:0", our_classes())
obs = [torch.from_numpy(np.random.default_rng(i).uniform(-1, 1, (E, 1, w["obs_dim"])).astype(np.float32)) for i in range(4)]
for i in range(200):
    model(cond={"state": obs[i % 4]})
best = []
for rep in range(7):
    t0 = time.perf_counter()
    for i in range(2000):
        model(cond={"state": obs[i % 4]})
    best.append((time.perf_counter() - t0) / 2000 * 1e6)
print("DPPO_B200_TICKET=%s  us per decision: min %.2f  median %.2f  all %s" % (
    os.environ.get("DPPO_B200_TICKET", "1"), min(best), sorted(best)[3], " ".join("%.1f" % b for b in best)))
