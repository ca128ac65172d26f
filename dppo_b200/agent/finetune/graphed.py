"""
CUDA-graph replay of one PPO minibatch (gradient zeroing, actor_ft / critic forward, fused loss kernel, autograd
backward, gradient all-reduce): ~150 launches replayed with one host call.

The update step of the reference loop (/root/reference/dppo/agent/finetune/train_ppo_diffusion_agent.py:313-364) is
launch-bound once its GEMMs run on the tensor cores: measured on cfg2, 3.4 ms of wall time per 50 000-row minibatch for
2.7 ms of GPU work on one GPU, and no speed-up at all from a second GPU (each rank's 25 000 rows are 1.4 ms of GPU work
behind the same 3 ms of Python / launch overhead).  The minibatch indices are the only per-call input: they are copied
into a static buffer, everything else (rollout buffers, parameters, the flat gradient buffer) keeps its address for
the lifetime of the graph.  The optimiser steps and the one device->host read of the diagnostics stay outside.
"""

import torch


class GraphedMinibatch:
    """`fn(inds)` must be free of host synchronisation and use only tensors whose storage outlives the graph."""

    def __init__(self, fn, n_inds, device, warmup=3, after_replay=None):
        self.fn = fn
        self.after_replay = after_replay  # host-side bookkeeping the captured work implies (e.g. version counters)
        self.static_inds = torch.zeros(n_inds, dtype=torch.int64, device=device)
        self.graph = None
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(self.static_inds)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn(self.static_inds)
        self.graph = g

    def __call__(self, inds):
        self.static_inds.copy_(inds)
        self.graph.replay()
        if self.after_replay is not None:
            self.after_replay()


class MinibatchStep:
    """
    One whole PPO minibatch as a stream-ordered unit without host synchronisation:

        fwd_bwd(inds)  (gradient zeroing, forward, fused loss, backward, gradient all-reduce)
        AdamW(actor_ft) [skipped during the critic warm-up], AdamW(critic)      (dppo_adamw_flat_dev, step count on the device)
        dppo_kl_check   (diagnostics of minibatch k -> history[k]; raises the device stop flag when approx_kl > target_kl)

    optionally captured once and replayed as ONE CUDA graph.  Once the stop flag is up, later optimiser launches do nothing,
    so what is applied equals the reference's loop with its host-side `break` (train_ppo_diffusion_agent.py:313-382) while
    the host only looks at the flag with a lag.
    """

    def __init__(self, fwd_bwd, grads, actor_opt, critic_opt, with_actor, max_grad_norm, target_kl, n_max, batch_size, device,
                 use_graph=True, world=1):
        from dppo_b200 import _lib

        self._lib, self.lib = _lib, _lib.load()
        self.fwd_bwd, self.grads = fwd_bwd, grads
        self.actor_opt, self.critic_opt, self.with_actor, self.max_grad_norm = actor_opt, critic_opt, with_actor, max_grad_norm
        self.target_kl, self.n_max = target_kl, int(n_max)
        self.kl_state = torch.zeros(4, dtype=torch.int32, device=device)  # [stop, index of the stopping minibatch, counter]
        self.history = torch.zeros((self.n_max, 8), dtype=torch.float32, device=device)
        self.stop_flag = self.kl_state[0:1]
        self.graphed = None
        self.graph_error = None
        if use_graph:
            ok = 1
            self.kl_state[0] = 1  # warm-up / capture launches must not apply optimiser steps
            try:
                self.graphed = GraphedMinibatch(self._step, batch_size, device)
            except Exception as ex:  # noqa: BLE001 - capture is an optimisation only
                self.graph_error = f"{type(ex).__name__}: {str(ex)[:200]}"
                torch.cuda.synchronize(device)
                ok = 0
            if world > 1:  # every rank replays or every rank launches eagerly (the all-reduce is inside)
                import torch.distributed as dist

                flag = torch.tensor([ok], dtype=torch.int32, device=device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                ok = int(flag.item())
            if not ok:
                self.graphed = None
            self.kl_state.zero_()
            self.history.zero_()

    def _step(self, inds):
        L = self._lib
        out = self.fwd_bwd(inds)
        if self.with_actor:
            self.actor_opt.step(max_grad_norm=self.max_grad_norm, stop_flag=self.stop_flag, bump=False)
        self.critic_opt.step(stop_flag=self.stop_flag, bump=False)
        L.check(self.lib.dppo_kl_check(L.ptr(self.grads.scalars), float(self.target_kl or 0.0), int(self.target_kl is not None),
                                       L.ptr(self.kl_state), L.ptr(self.history), self.n_max, L.stream_ptr()), "dppo_kl_check")
        return out

    def __call__(self, inds):
        if self.graphed is not None:
            self.graphed(inds)
            return None
        return self._step(inds)

    def finish(self):
        """After the last launch: mark the parameters as changed for the packed-weight caches of the chain kernels."""
        if self.with_actor:
            self.actor_opt.bump_versions()
        self.critic_opt.bump_versions()
