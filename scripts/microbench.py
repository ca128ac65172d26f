"""tcgen05.mma issue rate vs N and L2 -> shared-memory streaming rate (plain / cluster multicast).  Bring-up aid."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from dppo_b200 import _lib

lib = _lib.load()
lib.dppo_debug_mma_rate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
lib.dppo_debug_stream_rate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
out = torch.zeros(256, dtype=torch.int64, device="cuda")
n = 8192
print("== tcgen05.mma M=128 K=16 bf16 SS: cycles per MMA (tensor floor N/2, smem floor (4096+32N)/128)")
for N in (16, 32, 48, 64, 96, 128, 256):
    for n_b in (1, 2):
        if n_b * N > 512:
            continue
        for _ in range(2):
            rc = lib.dppo_debug_mma_rate(N, n, n_b, C.c_void_p(out.data_ptr()), None)
            torch.cuda.synchronize()
        print(f"N={N:3d} n_b={n_b} rc={rc} cycles/MMA {out[0].item() / n:7.1f}   floors tensor {N / 2:5.1f} smem {(4096 + 32 * N) / 128:5.1f}")

print("== L2 -> smem streaming of one shared region, 16 KiB tiles, 8-stage ring: bytes/cycle/SM")
region_tiles = 140  # ~2.2 MiB, the hi+lo tile stream of one 512-wide denoiser step
region = torch.zeros(region_tiles * 16384, dtype=torch.uint8, device="cuda")
n_tiles = 4000
for grid in (8, 32, 64, 96, 128, 144):
    for cl in (1, 2, 4, 8):
        if grid % cl:
            continue
        for _ in range(2):
            out.zero_()
            rc = lib.dppo_debug_stream_rate(C.c_void_p(region.data_ptr()), region_tiles, n_tiles, grid, cl, C.c_void_p(out.data_ptr()), None)
            torch.cuda.synchronize()
        cyc = out[:grid].double()
        bpc = n_tiles * 16384 / cyc.max().item()
        print(f"grid={grid:3d} cluster={cl} rc={rc}  {bpc:6.1f} B/cyc/SM   chip {bpc * grid:8.0f} B/cyc   (max {cyc.max().item() / 1e3:.0f} kcyc, min {cyc.min().item() / 1e3:.0f})")
