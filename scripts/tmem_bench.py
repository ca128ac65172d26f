"""tcgen05.ld read-back rate of a 128 x 256 fp32 accumulator by 8 warps: load shapes / batching, with and without a
concurrent N = 256 MMA stream.  Bring-up aid.  python scripts/tmem_bench.py"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from dppo_b200 import _lib

lib = _lib.load_test()
lib.dppo_debug_tmem_read_rate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
out = torch.zeros(16, dtype=torch.int64, device="cuda")
names = {0: "x16 + wait each", 1: "x32 + wait each", 2: "2 x x32, one wait", 3: "4 x x32, one wait", 4: "4 x x16, one wait"}
iters = 200
for mma in (0, 1):
    for v in range(5):
        for _ in range(2):
            out.zero_()
            rc = lib.dppo_debug_tmem_read_rate(v, iters, mma, C.c_void_p(out.data_ptr()), None)
            torch.cuda.synchronize()
        cyc = out[:8].max().item() / iters
        extra = f"  MMA {out[8].item() / max(1, out[9].item()):6.1f} cycles each" if mma else ""
        print(f"mma={mma} {names[v]:20s} rc={rc}  {cyc:8.1f} cycles per 128x256 tile  ({128 * 256 * 4 / cyc:6.1f} B/cycle){extra}")
