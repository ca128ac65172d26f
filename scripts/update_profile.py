"""Where one PPO minibatch goes (cfg2, 50 000 rows): GPU-busy time vs wall time, top kernels (torch.profiler)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from dppo_b200 import distributed as D
from dppo_b200.optim import FlatAdamW
from dppo_b200.workloads import get_workload
from tests.helpers import build_model, our_classes

name = sys.argv[1] if len(sys.argv) > 1 else "walker2d"
w = get_workload(name); dev = torch.device("cuda:0"); E = w["n_envs"]
model = build_model(w, "cuda:0", our_classes())
ft, Ta, Da = w["ft_denoising_steps"], w["horizon_steps"], w["action_dim"]
n_steps = max(1, min(w["train"]["n_steps"], (1 << 21) // max(1, E * ft))); N = n_steps * E
g = torch.Generator(device=dev).manual_seed(7)
obs_k = torch.rand((N, w["cond_steps"], w["obs_dim"]), device=dev, generator=g) * 2 - 1
chains_k = torch.randn((N, ft + 1, Ta, Da), device=dev, generator=g).clamp_(-1, 1)
with torch.no_grad():
    logprobs_k = torch.empty((N, ft, Ta, Da), device=dev)
    for s in range(0, N, 32768):
        logprobs_k[s:s + 32768] = model.get_logprobs({"state": obs_k[s:s + 32768]}, chains_k[s:s + 32768]).view(-1, ft, Ta, Da)
    values_k = model.critic({"state": obs_k}).view(-1)
adv_k = torch.randn(N, device=dev, generator=g); ret_k = adv_k + values_k
bs = min(w["train"]["batch_size"], N * ft)
opt_a = FlatAdamW(model.actor_ft.parameters(), lr=1e-4, weight_decay=0); opt_c = FlatAdamW(model.critic.parameters(), lr=1e-3, weight_decay=0)
grads = D.FlatGradBuffer([list(model.actor_ft.parameters()), list(model.critic.parameters())])
def minibatch(perm=True):
    inds = D.broadcast_permutation(N * ft, dev)[:bs] if perm else fixed
    grads.zero()
    res = model.loss_gathered(obs_k, chains_k, logprobs_k, ret_k, values_k, adv_k, inds, row_begin=0, row_count=bs,
                              reward_horizon=w["act_steps"], scalars_out=grads.scalars)
    (res[0] + 0.5 * res[2]).backward()
    kl = grads.scalars.tolist()[2]
    opt_a.step(); opt_c.step()
fixed = torch.randperm(N * ft, device=dev)[:bs]
for _ in range(3): minibatch()
for label, perm in (("with randperm", True), ("fixed indices", False)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): minibatch(perm)
    b.record(); torch.cuda.synchronize()
    print(f"{label}: wall {(time.perf_counter() - t0) * 100:.2f} ms / minibatch, events {a.elapsed_time(b) / 10:.2f} ms")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): minibatch(False)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))
