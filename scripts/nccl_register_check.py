"""NCCL user-buffer registration of the flat gradient buffer: all-reduce latency registered vs plain, eager and inside a
CUDA graph, and a clean exit.  torchrun --nproc-per-node N scripts/nccl_register_check.py [release|norelease] [MB]"""
import faulthandler
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

faulthandler.dump_traceback_later(70, exit=True)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
mode = sys.argv[1] if len(sys.argv) > 1 else "release"
mb = float(sys.argv[2]) if len(sys.argv) > 2 else 28.8
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
from dppo_b200 import distributed as D  # noqa: E402

n = int(mb * 1e6 / 4)


def timed(buf, graph):
    def f():
        dist.all_reduce(buf)

    for _ in range(5):
        f()
    torch.cuda.synchronize()
    g = None
    if graph:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            f()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            f()
        f = g.replay
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        f()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / 50], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del g
    return float(t) * 1e3


plain = torch.zeros(n, device=dev)
reg, how = D.alloc_collective_buffer(n, dev)
for name, buf in (("plain", plain), ("registered", reg)):
    for graph in (False, True):
        us = timed(buf, graph)
        if rank == 0:
            print(f"{name:10s} graph={int(graph)}  {us:8.1f} us  bus {2 * (world - 1) / world * n * 4 / us / 1e3:7.1f} GB/s   ({how if name != 'plain' else 'cudaMalloc'})",
                  flush=True)
t0 = time.time()
if mode == "release":
    del reg, plain
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
if rank == 0:
    print(f"exit path '{mode}' done in {time.time() - t0:.2f} s", flush=True)
faulthandler.cancel_dump_traceback_later()
