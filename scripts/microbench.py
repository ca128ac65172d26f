"""tcgen05.mma issue rate vs N and L2 -> shared-memory streaming rate (plain / cluster multicast).  Bring-up aid."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from dppo_b200 import _lib

lib = _lib.load_test()
lib.dppo_debug_mma_rate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
lib.dppo_debug_stream_rate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
out = torch.zeros(256, dtype=torch.int64, device="cuda")
n = 8192
print("== tcgen05.mma M=128 K=16 bf16 SS: cycles per MMA (tensor floor N/2, smem floor (4096+32N)/128)")
for N in (16, 32, 48, 64, 96, 128, 256):
    for n_b in (1, 2):
        if n_b * N > 256:
            continue
        for _ in range(2):
            rc = lib.dppo_debug_mma_rate(N, n, n_b, 0, 0, None, C.c_void_p(out.data_ptr()), None)
            torch.cuda.synchronize()
        plain = out[0].item() / n
        for _ in range(2):
            rc2 = lib.dppo_debug_mma_rate(N, n, n_b, 1, 0, None, C.c_void_p(out.data_ptr()), None)
            torch.cuda.synchronize()
        print(f"N={N:3d} n_b={n_b} rc={rc},{rc2} cycles/MMA {plain:7.1f}   floors tensor {N / 2:5.1f} smem {(4096 + 32 * N) / 128:5.1f}"
              f"   | pairs sharing A through the collector (fill / lastuse): {out[0].item() / n:7.1f}  smem floor {(2048 + 32 * N) / 128:5.1f}")

print("== the same N=64 stream next to what the chain kernel runs concurrently (1 = 16 KiB bulk ingest x4 in flight, 2 = tcgen05.ld loop in 2 warps, 4 = st.shared loop in 2 warps)")
ingest_src = torch.zeros(65536, dtype=torch.uint8, device="cuda")
for N in (64, 128):
    for load in (0, 1, 2, 4, 3, 5, 7):
        for _ in range(2):
            out.zero_()
            rc = lib.dppo_debug_mma_rate(N, n, 2, 0, load, C.c_void_p(ingest_src.data_ptr()), C.c_void_p(out.data_ptr()), None)
            torch.cuda.synchronize()
        cyc = out[0].item()
        print(f"N={N:3d} load={load} rc={rc} cycles/MMA {cyc / n:7.1f}   ingest {out[1].item() / cyc:5.1f} B/cyc   side-loop iterations/kcyc {out[2].item() * 1e3 / cyc:7.1f}")

print("== tcgen05.mma.cta_group::2 M=256 K=16 bf16 SS over a CTA pair: cycles per MMA, layout check of the accumulator")
lib.dppo_debug_mma_pair_rate.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
for N in (64, 128, 256):
    d = torch.zeros(2, 128, N, device="cuda")
    exp = torch.zeros(2, 128, N)
    for _ in range(2):
        rc = lib.dppo_debug_mma_pair_rate(N, n, C.c_void_p(d.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(exp.data_ptr()), None)
        torch.cuda.synchronize()
    ok = torch.equal(d.cpu(), exp)
    print(f"N={N:3d} (N/2={N // 2} rows of B per CTA) rc={rc} layout {'ok' if ok else 'MISMATCH'} cycles/MMA {out[0].item() / n:7.1f}"
          f"   floors per SM: tensor {N / 2:5.1f} smem {(4096 + 16 * N) / 128:5.1f}")
    if not ok:
        bad = (d.cpu() != exp).nonzero()
        print("   first mismatches", bad[:4].tolist(), "got", d.cpu()[tuple(bad[0])].item(), "want", exp[tuple(bad[0])].item())

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
