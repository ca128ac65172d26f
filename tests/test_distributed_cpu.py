"""
N > 1 host logic on CPU: world_size-2 (and -3, uneven) gloo process groups exercise dppo_b200/distributed.py —
env sharding, the once-per-iteration all-gather into the reference's (step, env) order, the broadcast permutation and
per-rank minibatch slices (bit-exact union), and the flat gradient + diagnostics all-reduce.
"""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dppo_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(world, fn):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    return [ret[r] for r in range(world)]


def test_shards_partition_envs_and_rows():
    for E in (1, 40, 50, 1000, 4096):
        for W in (1, 2, 3, 4, 8):
            spans = [D.env_shard(E, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == E
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1
    for B in (7, 10000, 17600, 50000):
        for W in (1, 2, 4, 8):
            spans = [D.minibatch_slice(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _gather_case(rank, world):
    n_steps, E = 5, 7  # uneven over 2 or 3 ranks
    full = torch.arange(n_steps * E * 3, dtype=torch.float32).view(n_steps, E, 3)
    b, e = D.env_shard(E, rank, world)
    got = D.gather_env_dim(full[:, b:e].contiguous(), E)
    flat_ok = torch.equal(got.view(n_steps * E, 3), full.view(n_steps * E, 3))  # row = step * E + env, as the reference
    vec = torch.arange(n_steps * E, dtype=torch.float64).view(n_steps, E)
    got2 = D.gather_env_dim(vec[:, b:e].contiguous(), E)
    return bool(flat_ok and torch.equal(got2, vec))


@pytest.mark.parametrize("world", [2, 3])
def test_gather_restores_reference_row_order(world):
    assert all(_run(world, _gather_case))


def _perm_case(rank, world):
    torch.manual_seed(1234 + rank)  # different generator state per rank: the broadcast must win
    total, bs, ft = 97, 40, 5
    perm = D.broadcast_permutation(total, "cpu")
    rows = []
    for batch in range(max(1, total // bs)):
        inds_b = perm[batch * bs:(batch + 1) * bs]
        lo, hi = D.minibatch_slice(inds_b.numel(), rank, world)
        mine = inds_b[lo:hi]
        rows.append((mine // ft).tolist() + (mine % ft).tolist())
    return perm.tolist(), rows


def test_minibatch_slices_union_is_the_reference_minibatch():
    world = 2
    out = _run(world, _perm_case)
    perm0 = out[0][0]
    assert all(o[0] == perm0 for o in out) and sorted(perm0) == list(range(97))
    total, bs, ft = 97, 40, 5
    for batch in range(total // bs):
        ref = torch.tensor(perm0[batch * bs:(batch + 1) * bs])
        ref_b, ref_d = (ref // ft).tolist(), (ref % ft).tolist()  # torch.unravel_index(inds, (N, ft))
        got_b, got_d = [], []
        for r in range(world):
            row = out[r][1][batch]
            got_b += row[: len(row) // 2]
            got_d += row[len(row) // 2:]
        assert got_b == ref_b and got_d == ref_d  # bit-exact integer indexing, order preserved


def _grad_case(rank, world):
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    x, y = torch.randn(32, 6), torch.randn(32, 1)
    buf = D.FlatGradBuffer(list(net.parameters()), n_scalars=8)
    lo, hi = D.minibatch_slice(32, rank, world)
    buf.zero()
    loss = ((net(x[lo:hi]) - y[lo:hi]) ** 2).sum() / 32  # partial sum already divided by the GLOBAL row count
    loss.backward()
    buf.scalars[0] = loss.detach()
    aliased = all(p.grad.data_ptr() >= buf.flat.data_ptr() for p in net.parameters())
    buf.allreduce()
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    ref.load_state_dict(net.state_dict())
    full = ((ref(x) - y) ** 2).mean()
    full.backward()
    err = max(float((p.grad - q.grad).abs().max()) for p, q in zip(net.parameters(), ref.parameters()))
    return aliased, err, abs(float(buf.scalars[0]) - float(full))


def test_flat_gradient_allreduce_equals_full_batch():
    for aliased, err, lerr in _run(2, _grad_case):
        assert aliased and err < 1e-6 and lerr < 1e-6


def _split_case(rank, world):
    """The two collectives of FlatGradBuffer.allreduce_split (group 0, then the rest + scalars) cover the buffer exactly once."""
    import torch.distributed as dist

    torch.manual_seed(1)
    a, c = torch.nn.Linear(7, 3), torch.nn.Linear(3, 1)  # 24 / 4 parameters: both groups need their 16-byte padding
    buf = D.FlatGradBuffer([list(a.parameters()), list(c.parameters())], n_scalars=8)
    g = torch.Generator().manual_seed(10 + rank)
    buf.flat.copy_(torch.randn(buf.flat.numel(), generator=g))
    whole = buf.flat.clone()
    dist.all_reduce(whole)
    setup = buf.overlap_setup()  # CPU / gloo: no side stream, the callers keep the single all-reduce
    dist.all_reduce(buf.segment(0))
    dist.all_reduce(buf.flat[buf.group_sizes[0]:])
    views_ok = (a.weight.grad.data_ptr() == buf.segment(0).data_ptr()
                and c.weight.grad.data_ptr() == buf.segment(1).data_ptr()
                and buf.scalars.data_ptr() == buf.flat[sum(buf.group_sizes):].data_ptr())
    return setup is None, views_ok, bool(torch.equal(buf.flat, whole))


def test_split_allreduce_pieces_equal_the_single_allreduce():
    for no_setup, views_ok, same in _run(2, _split_case):
        assert no_setup and views_ok and same
