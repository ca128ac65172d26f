#!/usr/bin/env python
"""
bench.py — DPPO hot-path benchmark (contract: see the repo's task statement; design notes in DESIGN.md §Measurement).

One "step" = one rollout decision for every environment of the workload: the S-step denoising chain
(VPGDiffusion.forward with return_chain=True) over E synthetic observations.  Metric = policy env-steps/s
= E * act_steps / t_step, the reference's own step accounting (train_ppo_diffusion_agent.py:151).

  value      chain kernel with inputs resident in HBM (CUDA events around each step, L2 flushed between steps)
  e2e        the same through the reference-shaped call model(cond=...) with HOST observations in and trajectories +
             chains in host memory on return, every step (what the agent does at train_ppo_diffusion_agent.py:107-122;
             dppo_sample_chain_host: the kernel moves both directions over PCIe itself, the call synchronises);
             e2e_explicit_copies is the same with copy launches around the device call
  update     PPO-update samples/s: fused gather+log-prob+loss kernel, autograd backward, optimiser steps (secondary)
  roofline   tensor-core roofline of the chain kernel from the algorithmic FLOPs S * F_net per decision
  cpu_baseline / --impl reference   the CPU oracle port of the reference path on this box's host cores
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from dppo_b200.workloads import chain_evals, get_workload  # noqa: E402

METRIC = "policy env-steps/sec (K-step denoise)"
UNIT = "env-steps/s"


def net_flops_per_sample(w):
    """Forward FLOPs of one denoiser evaluation (2 * MACs of every Linear), SURVEY.md §8 F_net."""
    a = w["actor"]
    if a["kind"] == "unet":
        # torch.utils.flop_counter on the reference Unet1D of cfg5 (F_net = 5,359,104; convs counted
        # as written, i.e. including the taps that fall into the zero padding)
        if (a["dim"], tuple(a["dim_mults"]), a["kernel_size"], w["horizon_steps"], w["action_dim"]) != (64, (1, 2), 5, 4, 7):
            raise NotImplementedError("F_net is tabulated for the cfg5 Unet1D only")
        return 5359104
    D, cond = w["horizon_steps"] * w["action_dim"], w["obs_dim"] * w["cond_steps"]
    td, dims = a["time_dim"], a["mlp_dims"]
    macs = td * 2 * td + 2 * td * td
    c_out = cond
    if a.get("cond_mlp_dims"):
        c0, c1 = a["cond_mlp_dims"]
        macs += cond * c0 + c0 * c1
        c_out = c1
    H = dims[0]
    macs += (D + td + c_out) * H + (len(dims) - 1) * H * H + H * D
    return 2 * macs


def workload_name(w, E):
    kind = f"DDIM-{w['ddim_steps']}" if w["use_ddim"] else f"DDPM K={w['denoising_steps']}"
    net = "unet" if w["actor"]["kind"] == "unet" else "mlp"
    return (f"{w['yaml'].split('/')[2]} ft_ppo_diffusion_{net}, {E} synthetic envs, obs {w['obs_dim']}, act {w['action_dim']}, "
            f"Tp=Ta={w['horizon_steps']}, {kind}, ft={w['ft_denoising_steps']}")


def bench_config(w, E):
    """The `config` object of BOTH arms (identical dicts): the workload the metric is quoted on."""
    return {"workload": workload_name(w, E)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------- CPU arm
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_classes():
    """The UNMODIFIED reference classes from baseline/_ref/dppo (a git-ignored copy of /root/reference/dppo that travels
    with the working tree), or None when it is absent / not importable."""
    if not os.path.isdir(os.path.join(REF_DIR, "dppo")):
        return None
    try:
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        from dppo.model.common.critic import CriticObs
        from dppo.model.diffusion.diffusion_ppo import PPODiffusion
        from dppo.model.diffusion.eta import EtaFixed
        from dppo.model.diffusion.mlp_diffusion import DiffusionMLP
        from dppo.model.diffusion.unet import Unet1D

        return dict(ppo=PPODiffusion, mlp=DiffusionMLP, unet=Unet1D, critic=CriticObs, eta=EtaFixed)
    except Exception as ex:  # noqa: BLE001 - any import problem means "use the port"
        print(f"[bench] reference classes not importable ({type(ex).__name__}: {ex}); timing the oracle port", file=sys.stderr)
        return None


def cpu_chain_seconds(w, E, reps, warmup, threads):
    """(times, kind): the reference's own VPGDiffusion.forward (model(cond, deterministic=False, return_chain=True), CPU
    fp32, all host threads) when baseline/_ref is present -> kind "reference"; else the oracle port of the same path (both
    networks on fine-tuned steps, like the reference) -> kind "port"."""
    from tests.helpers import build_model, make_inputs, oracle_cfgs, oracle_params, our_classes

    torch.set_num_threads(threads)
    ref = reference_classes()
    if ref is not None:
        model = build_model(w, "cpu", ref)
        model.train()
        state = make_inputs(w, E, 8)["state"]
        times = []
        with torch.no_grad():
            for i in range(warmup + reps):
                t0 = time.perf_counter()
                model(cond={"state": state}, deterministic=False, return_chain=True)
                if i >= warmup:
                    times.append(time.perf_counter() - t0)
        return times, "reference"
    from oracle import dppo_oracle as O

    model = build_model(w, "cpu", our_classes())
    nc, dc = oracle_cfgs(w)
    p = oracle_params(model)
    inp = make_inputs(w, E, 8)
    times = []
    for i in range(warmup + reps):
        t0 = time.perf_counter()
        O.sample_chain(p, nc, dc, inp["state"], inp["noise"], faithful_cost=True)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times, "port"


def cpu_kind_text(kind):
    return ("the reference's own PPODiffusion.forward (baseline/_ref/dppo, unmodified)" if kind == "reference"
            else "oracle port of the reference path (baseline/_ref absent)")


def run_reference(args, w, E, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    times, kind = cpu_chain_seconds(w, E, args.steps, args.warmup, cores)
    t = sum(times) / len(times)
    value = E * w["act_steps"] / t
    sample = f"{args.steps} full chains over all {E} envs ({cpu_kind_text(kind)}, torch CPU fp32, {cores} threads)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": bench_config(w, E),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- GPU arm
def run_b200(args, w, E, rank, world, local_rank):
    import torch.distributed as dist

    from tests.helpers import build_model, our_classes

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = build_model(w, str(dev), our_classes())
    model.engine_precision = args.precision
    eng = model.engine()
    S, ft, D, Do = chain_evals(w), w["ft_denoising_steps"], w["horizon_steps"] * w["action_dim"], w["obs_dim"] * w["cond_steps"]
    rng = np.random.default_rng(1000 + rank)
    n_bufs = 4
    states = [torch.from_numpy(rng.uniform(-1, 1, (E, w["cond_steps"], w["obs_dim"])).astype(np.float32)).to(dev) for _ in range(n_bufs)]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    min_std = float(model.get_min_sampling_denoising_std())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def kernel_step(i):
        return eng.sample(states[i % n_bufs], noise=None, seed=42, offset=i + 1, env_offset=rank * E, min_sampling_std=min_std)

    # ---- value: kernel-resident throughput
    clocks = ClockSampler(local_rank)
    clocks.start()
    for i in range(args.warmup):
        flush.zero_()
        kernel_step(i)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the event pair)
        ev[i][0].record()
        kernel_step(i)
        ev[i][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)

    # ---- e2e: pinned host obs in, trajectories + chains out, through the reference-shaped call
    host_obs = [torch.from_numpy(rng.uniform(-1, 1, (E, w["cond_steps"], w["obs_dim"])).astype(np.float32)).pin_memory() for _ in range(n_bufs)]
    host_traj = torch.empty((E, w["horizon_steps"], w["action_dim"]), dtype=torch.float32).pin_memory()
    host_chain = torch.empty((E, ft + 1, w["horizon_steps"], w["action_dim"]), dtype=torch.float32).pin_memory()

    # the drop-in call with HOST buffers: observations in host memory in, trajectories + chains in host memory out when the
    # call returns (dppo_sample_chain_host behind VPGDiffusion.forward: the kernel prologue reads the page-locked
    # observations over PCIe, the kernel stores the results into page-locked memory, the call synchronises)
    def e2e_step(i):
        out = model(cond={"state": host_obs[i % n_bufs]}, deterministic=False, return_chain=True)
        return out.trajectories, out.chains

    for i in range(args.warmup):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    barrier()
    e2e_s = time.perf_counter() - t0
    traj_h, chain_h = e2e_step(0)
    if traj_h.is_cuda or chain_h.is_cuda or not bool(torch.isfinite(chain_h).all()) or float(chain_h.abs().max()) == 0.0:
        raise RuntimeError("e2e: the host call did not return host-resident results")

    # the same with explicit copy launches around the device call (what the reference agent's code does literally)
    def e2e_copies_step(i):
        obs = host_obs[i % n_bufs].to(dev, non_blocking=True)
        out = model(cond={"state": obs}, deterministic=False, return_chain=True)
        host_traj.copy_(out.trajectories, non_blocking=True)
        host_chain.copy_(out.chains, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the agent consumes the actions on the host every step

    for i in range(args.warmup):
        e2e_copies_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_copies_step(i)
    barrier()
    e2e_cp_s = time.perf_counter() - t0

    # the B200 agent keeps the rollout buffers on the device: only the action chunk returns to the host
    def e2e_resident_step(i):
        obs = host_obs[i % n_bufs].to(dev, non_blocking=True)
        out = model(cond={"state": obs}, deterministic=False, return_chain=True)
        host_traj.copy_(out.trajectories, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for i in range(args.warmup):
        e2e_resident_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_resident_step(i)
    barrier()
    e2e_res_s = time.perf_counter() - t0
    # zero-copy variant of the same call: observations read from pinned host memory by the kernel prologue, trajectories
    # and chains stored by the kernel straight into pinned host memory (PCIe writes overlap the chain)
    def e2e_zero_copy_step(i):
        model(cond={"state": host_obs[i % n_bufs]}, deterministic=False, return_chain=True, out_trajectories=host_traj,
              out_chains=host_chain)
        torch.cuda.current_stream().synchronize()

    for i in range(args.warmup):
        e2e_zero_copy_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_zero_copy_step(i)
    barrier()
    e2e_zc_s = time.perf_counter() - t0
    # the timed regions are tens of milliseconds, nvidia-smi samples every 100 ms: keep issuing the SAME launches (untimed)
    # until the sampler has seen the device under this load for at least a second
    t_load = time.perf_counter()
    i = 0
    while time.perf_counter() - t_load < 1.2:
        kernel_step(i)
        i += 1
        if i % 16 == 0:
            torch.cuda.synchronize(dev)
    torch.cuda.synchronize(dev)
    clk = clocks.stop()
    clk["window"] = "warm-up + timed steps + e2e steps + 1.2 s of the same launches (untimed continuation)"

    # ---- update (secondary): one PPO minibatch = fused loss kernel + autograd backward + both optimiser steps
    upd = bench_update(args, w, model, dev, E, rank, world) if args.update else None
    # ---- strong scaling of north_star's sharded config (Furniture one_leg_low: 1000 envs and 17 600-row minibatches split
    # over the ranks, gradients all-reduced): the path that HAS a collective
    strong = None
    if args.strong and args.workload != "furniture":
        del model, eng
        torch.cuda.empty_cache()
        strong = bench_strong(args, dev, rank, world)

    t = torch.tensor([total_ms, e2e_s, wall, e2e_res_s, e2e_zc_s, e2e_cp_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s, wall, e2e_res_s, e2e_zc_s, e2e_cp_s = t.tolist()
    if rank == 0:
        act = w["act_steps"]
        value = world * E * act * args.steps / (total_ms * 1e-3)
        e2e_value = world * E * act * args.steps / e2e_s
        fnet = net_flops_per_sample(w)
        flops_launch = float(S) * fnet * E
        k_ms = total_ms / args.steps
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("bf16_tflops", 1590.0))
        achieved = flops_launch / (k_ms * 1e-3) / 1e12
        traffic = None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch of the latest `ncu --set full` capture (profiles/)
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{args.workload}:{E}")
        except (OSError, ValueError):
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": k_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3" if args.precision == "split3" else "bf16", "data": "synthetic",
            "config": bench_config(w, E),
            "setup": {"envs_per_gpu": E, "precision": args.precision,
                      "l2": "flushed between timed steps (512 MiB memset outside the event pairs)",
                      "weights": "random init seed 42, actor_ft perturbed 1e-2", "noise": "in-kernel Philox4x32-10",
                      "scaling": "weak: every rank runs the workload's envs, no data-path collective in the rollout metric; "
                                 "the collective path is measured in `strong_scaling`"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": E * Do * 4,
                    "d2h_bytes_per_step": E * D * 4 * (ft + 2), "ms_per_step": 1e3 * e2e_s / args.steps,
                    "note": "model(cond={'state': host tensor}) -> dppo_sample_chain_host: host observations in, trajectories + chains in host memory when the call returns (the kernel moves them over PCIe itself; stream synchronised inside the call)"},
            "e2e_explicit_copies": {
                "value": world * E * act * args.steps / e2e_cp_s, "unit": UNIT, "h2d_bytes_per_step": E * Do * 4,
                "d2h_bytes_per_step": E * D * 4 * (ft + 2), "ms_per_step": 1e3 * e2e_cp_s / args.steps,
                "note": "obs.to(device) -> model(cond) -> two copy_ launches into pinned memory -> stream synchronise (round 1's e2e)"},
            "e2e_zero_copy": {
                "value": world * E * act * args.steps / e2e_zc_s, "unit": UNIT, "h2d_bytes_per_step": E * Do * 4,
                "d2h_bytes_per_step": E * D * 4 * (ft + 2), "ms_per_step": 1e3 * e2e_zc_s / args.steps,
                "note": "device-path call with caller-owned pinned tensors passed as out_trajectories / out_chains + a torch stream synchronise: the kernel reads the observations from and stores trajectories + chains into page-locked host memory itself (no copy launches)"},
            "e2e_device_resident_buffers": {
                "value": world * E * act * args.steps / e2e_res_s, "unit": UNIT, "h2d_bytes_per_step": E * Do * 4,
                "d2h_bytes_per_step": E * D * 4, "ms_per_step": 1e3 * e2e_res_s / args.steps,
                "note": "same call; chains stay in the device-resident rollout buffer (what dppo_b200's agent does), only the action chunk is copied back"},
            "gpu_launches": args.steps,
            "clocks": clk,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "chain_unet_kernel" if w["actor"]["kind"] == "unet" else "chain_mlp_kernel",
                         "note": ("algorithmic FLOPs S*F_net*E (1x); split3 issues 3 bf16 MMAs per logical MMA so frac <= 1/3; "
                                  "inside a layer the kernel runs at the per-SM L2->shared weight ingest (34.7 B/cycle/SM "
                                  "measured), between the 80 dependent layers of a launch at hand-off latency, see DESIGN.md "
                                  "sections 3 and 6b; traffic = ncu dram bytes per launch (bytes); "
                                  + ("peak = MEASURED_PEAKS.json bf16_tflops (burst, kernel timed alone), of measured"
                                     if peaks else "peak = fallback 1.59 PF, of fallback"))},
            "wall_s_timed_region": wall,
            "update": upd,
        }
        line["strong_scaling"] = strong
        if args.cpu_baseline and world == 1:  # rank 0 at N = 1 only: under torchrun the other ranks' host threads compete for the cores
            cores = os.cpu_count() or 1
            reps = max(3, min(10, int(15.0 / max(0.05, 0.16e-3 * E))))
            times, kind = cpu_chain_seconds(w, E, reps, 1, cores)
            tc = sum(times) / len(times)
            line["cpu_baseline"] = {"value": E * act / tc, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"{reps} full chains over {E} envs, {cpu_kind_text(kind)}, torch CPU fp32, {cores} threads"}
            if args.workload != "hopper":
                line["hopper_50x_target"] = hopper_ratio(args, dev, cores)
        print(json.dumps(line), flush=True)
    if world > 1:
        from dppo_b200 import distributed as D

        D.shutdown()  # registered gradient buffers first, then the process group


def hopper_ratio(args, dev, cores):
    """north_star's explicit target: >= 50x the reference host-CPU env-steps/s for the 20-step DDPM Hopper chain (40 envs)
    on one B200, measured in this run on both sides (end to end: pinned host obs in, host actions + chains out)."""
    from tests.helpers import build_model, our_classes

    w = get_workload("hopper")
    E = w["n_envs"]
    model = build_model(w, str(dev), our_classes())
    model.engine_precision = args.precision
    ft = w["ft_denoising_steps"]
    rng = np.random.default_rng(5)
    host_obs = [torch.from_numpy(rng.uniform(-1, 1, (E, w["cond_steps"], w["obs_dim"])).astype(np.float32)).pin_memory() for _ in range(4)]
    host_traj = torch.empty((E, w["horizon_steps"], w["action_dim"]), dtype=torch.float32).pin_memory()
    host_chain = torch.empty((E, ft + 1, w["horizon_steps"], w["action_dim"]), dtype=torch.float32).pin_memory()

    def step(i):  # host observations in, host results out (dppo_sample_chain_host)
        model(cond={"state": host_obs[i % 4]}, deterministic=False, return_chain=True)

    def step_copies(i):
        obs = host_obs[i % 4].to(dev, non_blocking=True)
        out = model(cond={"state": obs}, deterministic=False, return_chain=True)
        host_traj.copy_(out.trajectories, non_blocking=True)
        host_chain.copy_(out.chains, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def step_zero_copy(i):
        model(cond={"state": host_obs[i % 4]}, deterministic=False, return_chain=True, out_trajectories=host_traj, out_chains=host_chain)
        torch.cuda.current_stream().synchronize()

    res = {}
    for name, fn in (("e2e", step), ("e2e_explicit_copies", step_copies), ("e2e_zero_copy", step_zero_copy)):
        for i in range(50):
            fn(i)
        blocks, n = [], 200  # median of five blocks of 200 decisions: one scheduler hiccup is 1-2 % of a 30 ms block
        for _ in range(5):
            t0 = time.perf_counter()
            for i in range(n):
                fn(i)
            blocks.append(time.perf_counter() - t0)
        res[name] = E * w["act_steps"] * n / statistics.median(blocks)
    times, kind = cpu_chain_seconds(w, E, 20, 3, cores)
    cpu = E * w["act_steps"] / (sum(times) / len(times))
    return {"workload": workload_name(w, E), "e2e_env_steps_s": res["e2e"], "e2e_explicit_copies_env_steps_s": res["e2e_explicit_copies"],
            "e2e_zero_copy_env_steps_s": res["e2e_zero_copy"],
            "cpu_env_steps_s": cpu, "cpu_kind": kind, "cores": cores, "ratio_e2e": res["e2e"] / cpu,
            "ratio_e2e_explicit_copies": res["e2e_explicit_copies"] / cpu,
            "ratio_e2e_zero_copy": res["e2e_zero_copy"] / cpu, "target": 50.0,
            "note": "e2e = model(cond={'state': host tensor}): host observations in, host trajectories + chains out, one call; GPU side = median of five blocks of 200 decisions, CPU side = mean of 20 chains"}


def bench_strong(args, dev, rank, world):
    """Furniture one_leg_low (reference cfg/furniture/finetune/one_leg_low/ft_ppo_diffusion_mlp.yaml:18-26,48,71): 1000
    envs and 17 600-row minibatches, SHARDED over the ranks (strong scaling), gradients of actor_ft + critic all-reduced
    every minibatch.  Reports rollout env-steps/s, update samples/s and the all-reduce alone."""
    import torch.distributed as dist

    from dppo_b200 import distributed as D
    from dppo_b200.optim import FlatAdamW
    from tests.helpers import build_model, our_classes

    w = get_workload("furniture")
    E_glob, bs_glob = w["n_envs"], w["train"]["batch_size"]
    e0, e1 = D.env_shard(E_glob, rank, world)
    E = e1 - e0
    model = build_model(w, str(dev), our_classes())
    model.engine_precision = args.precision
    eng = model.engine()
    ft, Ta, Da = w["ft_denoising_steps"], w["horizon_steps"], w["action_dim"]
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    state = torch.rand((E, w["cond_steps"], w["obs_dim"]), device=dev, generator=g) * 2 - 1
    min_std = float(model.get_min_sampling_denoising_std())

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_ms(fn, reps):
        for _ in range(3):
            fn()
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        sync_all()
        t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    chain_ms = timed_ms(lambda: eng.sample(state, seed=1, offset=1, env_offset=e0, min_sampling_std=min_std), 20)
    # rollout buffers of this rank's envs, gathered into the global (step, env) order like the agent does
    n_steps = 24
    obs_l = torch.rand((n_steps, E, w["cond_steps"], w["obs_dim"]), device=dev, generator=g) * 2 - 1
    with torch.no_grad():
        chains_l = torch.stack([model(cond={"state": obs_l[s]}, env_offset=e0).chains for s in range(n_steps)])
        lp_l = model.get_logprobs({"state": obs_l.flatten(0, 1)}, chains_l.flatten(0, 1)).view(n_steps, E, ft, Ta, Da)
        val_l = model.values({"state": obs_l.flatten(0, 1)}).view(n_steps, E)
    adv_l = torch.randn((n_steps, E), device=dev, generator=g)
    N = n_steps * E_glob
    obs_k = D.gather_env_dim(obs_l, E_glob).view(N, w["cond_steps"], w["obs_dim"]).contiguous()
    chains_k = D.gather_env_dim(chains_l, E_glob).view(N, ft + 1, Ta, Da).contiguous()
    lp_k = D.gather_env_dim(lp_l, E_glob).view(N, ft, Ta, Da).contiguous()
    val_k = D.gather_env_dim(val_l, E_glob).reshape(-1).contiguous()
    adv_k = D.gather_env_dim(adv_l, E_glob).reshape(-1).contiguous()
    ret_k = (adv_k + val_k).contiguous()
    opt_a = FlatAdamW(model.actor_ft.parameters(), lr=w["train"]["actor_lr"], weight_decay=0)
    opt_c = FlatAdamW(model.critic.parameters(), lr=w["train"]["critic_lr"], weight_decay=0)
    grads = D.FlatGradBuffer([list(model.actor_ft.parameters()), list(model.critic.parameters())])
    bs = min(bs_glob, N * ft)
    lo, hi = D.minibatch_slice(bs, rank, world)
    fused = model.fused_update_reason() is None
    overlap = grads.overlap_setup() if (fused and os.environ.get("DPPO_B200_OVERLAP", "1") == "1") else None

    def fwd_bwd(inds):
        grads.zero()
        if fused:
            model.update_minibatch(obs_k, chains_k, lp_k, ret_k, val_k, adv_k, inds, row_begin=lo, row_count=hi - lo,
                                   reward_horizon=w["act_steps"], vf_coef=w["train"]["vf_coef"], scalars_out=grads.scalars,
                                   actor_event=overlap[0] if overlap else None)
        else:
            res = model.loss_gathered(obs_k, chains_k, lp_k, ret_k, val_k, adv_k, inds, row_begin=lo, row_count=hi - lo,
                                      reward_horizon=w["act_steps"], scalars_out=grads.scalars)
            (res[0] + w["train"]["vf_coef"] * res[2]).backward()
        if overlap:  # actor segment next to the critic backward, the rest behind it (what the agent does)
            grads.allreduce_split()
        else:
            grads.allreduce()

    from dppo_b200.agent.finetune.graphed import MinibatchStep

    step_fn = MinibatchStep(fwd_bwd, grads, opt_a, opt_c, True, None, w["train"]["target_kl"], 256, bs, dev,
                            use_graph=args.graph_update, world=world)
    perm = D.broadcast_permutation(N * ft, dev)
    per_epoch = max(1, (N * ft) // bs)
    k = [0]

    def minibatch():
        j = k[0] % per_epoch
        k[0] += 1
        step_fn(perm[j * bs:(j + 1) * bs])

    upd_ms = timed_ms(minibatch, 10)
    ar_ms = timed_ms(grads.allreduce, 20) if world > 1 else 0.0
    nbytes = grads.flat.numel() * 4
    out = {
        "workload": workload_name(w, E_glob), "scaling": "strong", "n_gpus": world, "envs_per_gpu": E,
        "minibatch_rows_per_gpu": hi - lo, "minibatch_rows": bs,
        "rollout_env_steps_s": E_glob * w["act_steps"] / (chain_ms * 1e-3), "rollout_ms_per_decision": chain_ms,
        "update_samples_s": bs / (upd_ms * 1e-3), "update_ms_per_minibatch": upd_ms,
        "allreduce_bytes": nbytes, "allreduce_us": 1e3 * ar_ms,
        "allreduce_bus_GBps": (2.0 * (world - 1) / world * nbytes / (ar_ms * 1e-3) / 1e9) if world > 1 else None,
        "update_path": "dppo_update_minibatch (tcgen05 GEMM kernels)" if fused else "torch autograd",
        "gradient_buffer": grads.allocation, "cuda_graph": step_fn.graphed is not None,
        "note": "max over ranks of CUDA-event time; update = gather -> forward -> loss -> backward -> NCCL all-reduce of the flat "
                "gradient buffer -> both AdamW steps -> device-side KL check",
    }
    del model, eng, obs_k, chains_k, lp_k
    torch.cuda.empty_cache()
    return out


def bench_update(args, w, model, dev, E, rank, world):
    """PPO-update samples/s on synthetic rollout buffers produced by the sampler itself."""
    import torch.distributed as dist

    ft, Ta, Da = w["ft_denoising_steps"], w["horizon_steps"], w["action_dim"]
    n_steps = max(1, min(w["train"]["n_steps"], (1 << 21) // max(1, E * ft)))
    N = n_steps * E
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    obs_k = torch.rand((N, w["cond_steps"], w["obs_dim"]), device=dev, generator=g) * 2 - 1
    chains_k = torch.empty((N, ft + 1, Ta, Da), device=dev)
    with torch.no_grad():
        for s in range(n_steps):
            chains_k[s * E:(s + 1) * E] = model(cond={"state": obs_k[s * E:(s + 1) * E]}).chains
        logprobs_k = torch.empty((N, ft, Ta, Da), device=dev)
        chunk = 32768
        for s in range(0, N, chunk):
            logprobs_k[s:s + chunk] = model.get_logprobs({"state": obs_k[s:s + chunk]}, chains_k[s:s + chunk]).view(-1, ft, Ta, Da)
        values_k = model.values({"state": obs_k})
    # ---- once-per-iteration prologue (reference train_ppo_diffusion_agent.py:197-279): old log-probs of every stored
    # chain (dppo_chain_logprobs, teacher-forced chain kernel), running reward scaling + GAE (float64 scan kernels)
    from dppo_b200 import engine as E_
    from dppo_b200.util.reward_scaling import RunningRewardScalerCUDA

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / reps

    def all_logprobs():
        with torch.no_grad():
            for s0 in range(0, N, chunk):
                model.get_logprobs({"state": obs_k[s0:s0 + chunk]}, chains_k[s0:s0 + chunk])

    lp_ms = timed(all_logprobs)
    scaler = RunningRewardScalerCUDA(E, dev)
    rew = torch.randn((n_steps, E), dtype=torch.float64, device=dev, generator=g)
    firsts = (torch.rand((n_steps, E), device=dev, generator=g) < 0.01).double()
    term = (torch.rand((n_steps, E), device=dev, generator=g) < 0.01).double()
    vals = values_k.view(n_steps, E).double()
    nxt = torch.randn(E, dtype=torch.float64, device=dev, generator=g)
    gae_ms = timed(lambda: E_.gae(scaler(reward=rew, first=firsts), term, vals, nxt, 0.99, 0.95, 1.0), reps=10)
    prologue = {"logprob_rows_per_s": N * ft / (lp_ms * 1e-3), "logprob_ms": lp_ms, "rows": N * ft,
                "reward_scaling_plus_gae_ms": gae_ms, "gae_elements": n_steps * E,
                "gae_GBps_algorithmic": 20.0 * n_steps * E / (gae_ms * 1e-3) / 1e9,
                "path": "dppo_chain_logprobs (chain kernel, teacher-forced over the ft window); dppo_reward_scale_f64 + dppo_gae_f64"}
    adv_k = torch.randn(N, device=dev, generator=g)
    ret_k = adv_k + values_k
    from dppo_b200 import distributed as D

    bs = min(w["train"]["batch_size"], N * ft)
    from dppo_b200.optim import FlatAdamW

    opt_a = FlatAdamW(model.actor_ft.parameters(), lr=w["train"]["actor_lr"], weight_decay=0)
    opt_c = FlatAdamW(model.critic.parameters(), lr=w["train"]["critic_lr"], weight_decay=0)
    grads = D.FlatGradBuffer([list(model.actor_ft.parameters()), list(model.critic.parameters())])
    lo, hi = D.minibatch_slice(bs, rank, world)
    per_rank = bs // world

    fused = model.fused_update_reason() is None
    overlap = grads.overlap_setup() if (fused and os.environ.get("DPPO_B200_OVERLAP", "1") == "1") else None

    def fwd_bwd(inds):
        grads.zero()
        if fused:  # the whole minibatch inside libdppo_b200 (hand-written tcgen05 forward / dgrad / wgrad kernels)
            model.update_minibatch(obs_k, chains_k, logprobs_k, ret_k, values_k, adv_k, inds, row_begin=lo, row_count=hi - lo,
                                   reward_horizon=w["act_steps"], vf_coef=w["train"]["vf_coef"], with_actor=True,
                                   scalars_out=grads.scalars, actor_event=overlap[0] if overlap else None)
        else:
            res = model.loss_gathered(obs_k, chains_k, logprobs_k, ret_k, values_k, adv_k, inds, row_begin=lo,
                                      row_count=hi - lo, reward_horizon=w["act_steps"], scalars_out=grads.scalars)
            (res[0] + w["train"]["vf_coef"] * res[2]).backward()
        if overlap:  # multi-GPU: actor segment reduced next to the critic backward, the rest behind it
            grads.allreduce_split()
        else:
            grads.allreduce()  # gradients of both networks + loss diagnostics: one NCCL all-reduce

    # the agent's minibatch unit: forward / backward / all-reduce + both AdamW steps + device-side KL check, one CUDA graph
    from dppo_b200.agent.finetune.graphed import MinibatchStep

    step_fn = MinibatchStep(fwd_bwd, grads, opt_a, opt_c, True, None, w["train"]["target_kl"], 64, bs, dev,
                            use_graph=args.graph_update, world=world)
    graphed = step_fn.graphed
    if args.graph_update and graphed is None:
        print(f"[bench] CUDA-graph capture of the minibatch failed ({step_fn.graph_error}); running eagerly", file=sys.stderr, flush=True)
    perm = [None]

    def minibatch(k):
        # one permutation per epoch, sliced per minibatch, exactly like the reference loop
        # (train_ppo_diffusion_agent.py:311-316); an epoch of this buffer is N*ft // bs minibatches
        per_epoch = max(1, (N * ft) // bs)
        if perm[0] is None or k % per_epoch == 0:
            perm[0] = D.broadcast_permutation(N * ft, dev)
        j = k % per_epoch
        step_fn(perm[0][j * bs:(j + 1) * bs])

    for k in range(3):
        minibatch(k)
    torch.cuda.synchronize(dev)
    reps = 10
    perm[0] = None  # the timed region draws (and broadcasts) its own permutation
    t0 = time.perf_counter()
    for k in range(reps):
        minibatch(k)
    stop_state = step_fn.kl_state.tolist()  # the early-stop flag + the diagnostics history reach the host once, at the end
    torch.cuda.synchronize(dev)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    assert stop_state[0] == 0 and stop_state[2] == reps + 3, stop_state
    return {"metric": "PPO-update samples/sec", "value": per_rank * world * reps / float(dt), "unit": "samples/s",
            "minibatch_rows": per_rank * world, "buffer_rows": N * ft, "prologue_per_gpu": prologue,
            "path": ("fused gather+log-prob+loss fwd/bwd kernel; flat gradient buffer + one all-reduce; "
                     + ("actor_ft / critic forward, dgrad and wgrad as hand-written tcgen05 GEMM kernels on bf16 hi/lo operand images "
                        "(dppo_update_minibatch, 3-product split, fp32 accumulate), gradients accumulated straight into the flat buffer; "
                        if fused else
                        "Linear layers as 3-product bf16-split tensor-core GEMMs (dppo_split3_pack + cuBLASLt bf16, fp32 accumulate) under "
                        "torch autograd (" + str(model.fused_update_reason()) + "); ")
                     + "fused flat AdamW kernel per network (step count on the device), KL early-stop test on the device; "
                     + ("forward + loss + backward + all-reduce + optimiser steps + KL check replayed as ONE CUDA graph per minibatch"
                        if graphed else "eager launches")),
            "gradient_buffer": grads.allocation}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="walker2d")
    ap.add_argument("--envs", type=int, default=None, help="environments per GPU (default: the workload's n_envs)")
    ap.add_argument("--precision", default="split3", choices=["split3", "bf16"])
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-update", dest="update", action="store_false", help="skip the secondary PPO-update measurement")
    ap.add_argument("--no-strong", dest="strong", action="store_false",
                    help="skip the Furniture strong-scaling block (sharded envs + minibatches, gradient all-reduce)")
    ap.add_argument("--no-graph-update", dest="graph_update", action="store_false",
                    help="run the PPO minibatch eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    w = get_workload(args.workload)
    E = args.envs or w["n_envs"]
    if args.impl == "reference":
        run_reference(args, w, E, rank, world)
    else:
        run_b200(args, w, E, rank, world, local_rank)


if __name__ == "__main__":
    main()
