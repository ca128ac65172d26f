"""
Data-parallel plumbing of the DPPO hot path: one process per GPU, torch.distributed (NCCL on the B200 box, gloo in the
CPU tests).  The reference is single-process; this module is the only distributed component and it is
built so that W ranks reproduce the single-process arithmetic:

  rollout   envs are independent -> rank r owns the contiguous env range `env_shard(E, r, W)`; no collective while
            sampling.  Uneven shards are allowed (E = 50 over 4 ranks -> 13, 13, 12, 12).
  gather    once per iteration the rank-local rollout buffers (n_steps, E_r, ...) are all-gathered into the reference's
            (n_steps, E, ...) arrays (`gather_env_dim`), so the flat row index (step * E + env) * ft + d of
            train_ppo_diffusion_agent.py:316-320 is unchanged.
  update    rank 0 draws torch.randperm exactly like the reference (:311) and broadcasts it; rank r evaluates the
            contiguous slice `minibatch_slice(Bmb, r, W)` of every minibatch, divides by the GLOBAL row count, and
            `allreduce_flat` sums gradients and loss scalars in one collective.  The union of the slices is the
            reference minibatch bit-exactly; a remainder goes to the last rank.
"""

import logging
import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

log = logging.getLogger(__name__)
_LIVE_BUFFERS = []  # the FlatGradBuffers whose storage is registered with the NCCL communicator (until release())


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def env_shard(n_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[begin, end) of the envs rank `rank` owns: the first n_envs % W ranks get one extra env."""
    base, extra = divmod(n_envs, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def minibatch_slice(n_rows: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[begin, end) of the rows of one minibatch rank `rank` evaluates; the remainder goes to the last rank."""
    per = n_rows // world_size
    begin = rank * per
    return begin, (n_rows if rank == world_size - 1 else begin + per)


def gather_env_dim(local: torch.Tensor, n_envs: int, env_dim: int = 1) -> torch.Tensor:
    """
    All-gather rank-local buffers along their env dimension into the full array every rank then holds.
    `local` has size env_shard(...) along `env_dim`; shards are padded to the largest shard for the collective only.
    """
    rank, W = world()
    if W == 1:
        return local
    sizes = [env_shard(n_envs, r, W) for r in range(W)]
    widest = max(e - b for b, e in sizes)
    moved = local.movedim(env_dim, 0).contiguous()
    if moved.shape[0] < widest:
        pad = torch.zeros((widest - moved.shape[0],) + tuple(moved.shape[1:]), dtype=moved.dtype, device=moved.device)
        moved = torch.cat([moved, pad], 0)
    parts: List[torch.Tensor] = [torch.empty_like(moved) for _ in range(W)]
    dist.all_gather(parts, moved)
    full = torch.cat([p[: e - b] for p, (b, e) in zip(parts, sizes)], 0)
    return full.movedim(0, env_dim).contiguous()


def broadcast_permutation(n: int, device, generator=None) -> torch.Tensor:
    """torch.randperm(n) drawn on rank 0 (the reference's call, train_ppo_diffusion_agent.py:311) and broadcast."""
    rank, W = world()
    if W == 1:
        return torch.randperm(n, device=device, generator=generator)
    # the int64 permutation is drawn once (rank 0, the reference's generator stream); indices < 2^31 travel as int32
    wire_dtype = torch.int32 if n < 2 ** 31 else torch.int64
    if rank == 0:
        perm = torch.randperm(n, device=device, generator=generator)
        wire = perm.to(wire_dtype)
    else:
        perm = None
        wire = torch.empty(n, dtype=wire_dtype, device=device)
    dist.broadcast(wire, 0)
    return perm if rank == 0 else wire.to(torch.int64)


def alloc_collective_buffer(n: int, device) -> Tuple[torch.Tensor, str]:
    """Zeroed fp32 buffer of n elements for the per-minibatch gradient all-reduce.  Under NCCL it is allocated with
    ncclMemAlloc and registered with the communicator (ProcessGroupNCCL.allocate_tensor, or a torch MemPool on the
    backend's allocator + register_mem_pool), which makes it eligible for the zero-copy / NVLS (in-switch reduction) paths;
    anything else - gloo in the CPU tests, a single process, an older torch - gets a plain allocation.  Returns
    (tensor, how it was allocated).  DPPO_B200_NCCL_REGISTER=0 forces the plain path."""
    dev = torch.device(device)
    if (dev.type == "cuda" and dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl"
            and os.environ.get("DPPO_B200_NCCL_REGISTER", "1") == "1"):
        try:
            backend = dist.group.WORLD._get_backend(dev)
            if hasattr(backend, "supports_tensor_alloc") and backend.supports_tensor_alloc(dev):
                t = backend.allocate_tensor(n, dtype=torch.float32, device=dev)
                t.zero_()
                return t, "ncclMemAlloc + registered (ProcessGroupNCCL.allocate_tensor)"
            pool = torch.cuda.MemPool(backend.mem_allocator)
            with torch.cuda.use_mem_pool(pool):
                t = torch.zeros(n, dtype=torch.float32, device=dev)
            backend.register_mem_pool(pool)
            t._dppo_pool = pool  # keep the pool alive with the tensor
            return t, "ncclMemAlloc + ncclCommRegister (torch MemPool on the NCCL allocator)"
        except Exception as ex:  # noqa: BLE001 - registration is an optimisation
            log.warning("NCCL buffer registration unavailable (%s: %s); plain allocation", type(ex).__name__, ex)
    return torch.zeros(n, dtype=torch.float32, device=dev), "plain allocation"


class FlatGradBuffer:
    """
    One flat fp32 buffer aliasing the .grad of every trainable parameter, followed by `n_scalars` slots for loss
    diagnostics, so that one all-reduce per minibatch carries gradients and scalars.
    Backward writes straight into the views - there is no pack / unpack copy.
    """

    def __init__(self, params: Sequence, n_scalars: int = 8):
        """`params`: a list of parameters, or a list of parameter LISTS (one per optimiser): every group then starts on
        a 16-byte boundary, which is what the fused AdamW kernel (dppo_b200.optim.FlatAdamW) reads with float4 loads."""
        groups = [list(g) for g in params] if params and isinstance(params[0], (list, tuple)) else [list(params)]
        groups = [[p for p in g if p.requires_grad] for g in groups]
        self.params = [p for g in groups for p in g]
        n = sum((p.numel() + 3) // 4 * 4 for p in self.params)  # every tensor padded to 16 bytes (same layout as FlatAdamW)
        dev = self.params[0].device
        self.flat, self.allocation = alloc_collective_buffer(n + n_scalars, dev)
        if self.allocation != "plain allocation":
            _LIVE_BUFFERS.append(self)  # the .grad views keep the storage alive anyway; shutdown() must be able to find it
        self.n_grad, self.n_scalars = n, n_scalars
        self.group_sizes = [sum((p.numel() + 3) // 4 * 4 for p in g) for g in groups]
        o = 0
        for g in groups:
            for p in g:
                p.grad = self.flat[o:o + p.numel()].view_as(p)
                o += (p.numel() + 3) // 4 * 4

    @property
    def scalars(self) -> torch.Tensor:
        return self.flat[self.n_grad:]

    def zero(self):
        if self.flat.is_cuda:  # stream-ordered cudaMemsetAsync through the library: no framework fill kernel in the minibatch
            from dppo_b200 import _lib

            _lib.check(_lib.load().dppo_memset_zero(self.flat.data_ptr(), self.flat.numel() * 4, _lib.stream_ptr()), "dppo_memset_zero")
        else:
            self.flat.zero_()

    def allreduce(self):
        _, W = world()
        if W > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)

    def overlap_setup(self):
        """(event, stream) for allreduce_split: the event is handed to the update plan (UpdatePlan.set_actor_event), which
        records it behind the actor backward.  None on CPU / a single process."""
        _, W = world()
        if W <= 1 or not self.flat.is_cuda or len(self.group_sizes) < 2:
            return None
        if getattr(self, "_overlap", None) is None:
            ev = torch.cuda.Event()
            ev.record()  # creates the cudaEvent_t
            self._overlap = (ev, torch.cuda.Stream(device=self.flat.device))
        return self._overlap

    def allreduce_split(self):
        """Two collectives instead of one: group 0 (actor_ft) starts as soon as the recorded event says its gradients are
        final - issued from a side stream, so that it runs next to the critic backward still queued on the current stream -
        the rest (critic + loss diagnostics) behind the current stream's work.  Both go through the process group's NCCL
        stream in this order on every rank.  Captures into a CUDA graph as a fork / join."""
        ev, side = self._overlap
        cur = torch.cuda.current_stream()
        side.wait_event(ev)
        with torch.cuda.stream(side):
            dist.all_reduce(self.segment(0), op=dist.ReduceOp.SUM)
        dist.all_reduce(self.flat[self.group_sizes[0]:], op=dist.ReduceOp.SUM)
        cur.wait_stream(side)

    def release(self):
        """Drop the buffer (the parameters' .grad views included).  A buffer registered with the NCCL communicator must be
        gone before the process group is destroyed: destroy_process_group() with a live ncclMemAlloc'ed tensor hangs
        (observed on 2 x B200, torch 2.11 / NCCL 2.28) - shutdown() below does both in the right order."""
        for p in self.params:
            p.grad = None
        self.flat = None
        if self in _LIVE_BUFFERS:
            _LIVE_BUFFERS.remove(self)

    def segment(self, group: int) -> torch.Tensor:
        """The contiguous gradient segment of parameter group `group` (e.g. 0 = actor_ft, 1 = critic)."""
        o = sum(self.group_sizes[:group])
        return self.flat[o:o + self.group_sizes[group]]


def shutdown():
    """Release every NCCL-registered gradient buffer, then destroy the process group (no-op without one)."""
    import gc

    for buf in list(_LIVE_BUFFERS):
        buf.release()
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
