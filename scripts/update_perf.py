"""Time the PPO minibatch (dppo_update_minibatch + all-reduce-free single GPU) of a workload: eager launches with CUDA
events, then a CUDA-graph replay.  `--reps` small for an ncu launch list.

    python scripts/update_perf.py --workload walker2d --rows 50000 --reps 20
"""

import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dppo_b200 import distributed as D  # noqa: E402
from dppo_b200.optim import FlatAdamW  # noqa: E402
from dppo_b200.workloads import get_workload  # noqa: E402
from tests.helpers import build_model, our_classes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="walker2d")
    ap.add_argument("--rows", type=int, default=None)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--envs", type=int, default=None)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-actor", action="store_true")
    ap.add_argument("--n-steps", type=int, default=None, help="rollout steps in the synthetic buffer (small for ncu runs)")
    ap.add_argument("--profile-range", action="store_true",
                    help="cudaProfilerStart/Stop around the eager timed loop (ncu --profile-from-start off)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    w = get_workload(args.workload)
    model = build_model(w, str(dev), our_classes())
    ft, Ta, Da = w["ft_denoising_steps"], w["horizon_steps"], w["action_dim"]
    E = args.envs or w["n_envs"]
    bs = args.rows or w["train"]["batch_size"]
    n_steps = args.n_steps or max(1, min(w["train"]["n_steps"], (1 << 21) // max(1, E * ft)))
    N = n_steps * E
    g = torch.Generator(device=dev).manual_seed(7)
    obs_k = torch.rand((N, w["cond_steps"], w["obs_dim"]), device=dev, generator=g) * 2 - 1
    chains_k = torch.empty((N, ft + 1, Ta, Da), device=dev)
    with torch.no_grad():
        for s in range(n_steps):
            chains_k[s * E:(s + 1) * E] = model(cond={"state": obs_k[s * E:(s + 1) * E]}).chains
        logprobs_k = torch.empty((N, ft, Ta, Da), device=dev)
        for s in range(0, N, 32768):
            logprobs_k[s:s + 32768] = model.get_logprobs({"state": obs_k[s:s + 32768]}, chains_k[s:s + 32768]).view(-1, ft, Ta, Da)
        values_k = model.values({"state": obs_k})
    adv_k = torch.randn(N, device=dev, generator=g)
    ret_k = adv_k + values_k
    opt_a = FlatAdamW(model.actor_ft.parameters(), lr=w["train"]["actor_lr"], weight_decay=0)
    opt_c = FlatAdamW(model.critic.parameters(), lr=w["train"]["critic_lr"], weight_decay=0)
    grads = D.FlatGradBuffer([list(model.actor_ft.parameters()), list(model.critic.parameters())])
    bs = min(bs, N * ft)
    perm = torch.randperm(N * ft, device=dev)
    print(f"{args.workload}: buffer {N * ft} rows, minibatch {bs} rows, fused path: {model.fused_update_reason() is None}", flush=True)

    def fwd_bwd(inds):
        grads.zero()
        model.update_minibatch(obs_k, chains_k, logprobs_k, ret_k, values_k, adv_k, inds, reward_horizon=w["act_steps"],
                               vf_coef=w["train"]["vf_coef"], with_actor=not args.no_actor, scalars_out=grads.scalars)

    def step(fn, k):
        j = k % max(1, (N * ft) // bs)
        fn(perm[j * bs:(j + 1) * bs])
        opt_a.step()
        opt_c.step()

    for k in range(3):
        step(fwd_bwd, k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.profile_range:
        torch.cuda.cudart().cudaProfilerStart()
    t0 = time.perf_counter()
    a.record()
    for k in range(args.reps):
        step(fwd_bwd, k)
    b.record()
    torch.cuda.synchronize()
    if args.profile_range:
        torch.cuda.cudart().cudaProfilerStop()
    wall = (time.perf_counter() - t0) / args.reps
    ms = a.elapsed_time(b) / args.reps
    print(f"eager : {ms:.3f} ms / minibatch (device), {wall * 1e3:.3f} ms wall -> {bs / (ms * 1e-3) / 1e6:.2f} M samples/s", flush=True)
    if not args.no_graph:
        from dppo_b200.agent.finetune.graphed import GraphedMinibatch

        gm = GraphedMinibatch(fwd_bwd, bs, dev)
        for k in range(3):
            step(gm, k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a.record()
        for k in range(args.reps):
            step(gm, k)
        b.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / args.reps
        ms = a.elapsed_time(b) / args.reps
        print(f"graph : {ms:.3f} ms / minibatch (device), {wall * 1e3:.3f} ms wall -> {bs / (ms * 1e-3) / 1e6:.2f} M samples/s", flush=True)
    print("scalars", grads.scalars.tolist())


if __name__ == "__main__":
    main()
