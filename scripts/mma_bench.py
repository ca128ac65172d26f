import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from dppo_b200 import _lib

lib = _lib.load()
lib.dppo_debug_mma_bench.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_void_p]
out = torch.zeros(1, dtype=torch.int64, device="cuda")
n = 4096
print("N n_acc n_a per_commit -> cycles/MMA")
for N in (32, 64, 128, 256):
    for n_acc in (1, 2, 4):
        if n_acc * N > 512:
            continue
        for n_a in (1, 4, 8):
            for pc in (1 << 30, 8):
                for _ in range(2):
                    lib.dppo_debug_mma_bench(N, n_acc, n_a, n, pc, C.c_void_p(out.data_ptr()), None)
                    torch.cuda.synchronize()
                print(N, n_acc, n_a, "none" if pc > n else pc, "->", round(out.item() / n, 1))
