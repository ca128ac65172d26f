"""Three host-buffer decisions of the Hopper case (run under `ncu --metrics gpu__time_duration.sum` for the launch list)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dppo_b200.workloads import get_workload  # noqa: E402
from tests.helpers import build_model, our_classes  # noqa: E402

w = get_workload("hopper")
model = build_model(w, "cuda:0", our_classes())
obs = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, (w["n_envs"], 1, w["obs_dim"])).astype(np.float32))
model(cond={"state": obs})  # packs the weights (pack kernels) + first decision
torch.cuda.synchronize()
print("MARK: three decisions follow")
for _ in range(3):
    out = model(cond={"state": obs})
print("done", float(out.chains.abs().max()))
