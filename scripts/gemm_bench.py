"""Row-GEMM kernel sweep (bring-up): time per launch and role cycle counters for epilogue variants, tile widths and ring
depths.  python scripts/gemm_bench.py"""

import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dppo_b200 import _lib  # noqa: E402

lib = _lib.load()
lib.dppo_debug_gemm_bench.argtypes = [C.c_int] * 7 + [C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
lib.dppo_debug_gemm_bench.restype = C.c_int


def run(R, K, N, flags, ntile=0, stages=0, reps=10):
    prof = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    ms = C.c_float()
    rc = lib.dppo_debug_gemm_bench(R, K, N, flags, ntile, stages, reps, C.byref(ms), prof.data_ptr(), _lib.stream_ptr())
    torch.cuda.synchronize()
    if rc != 0:
        return f"rc={rc} {lib.dppo_last_error().decode()}"
    p = prof.view(148, 16).double()
    m = p.mean(0).tolist()
    flops = 2.0 * R * K * N
    return (f"{ms.value * 1e3:8.1f} us  {flops / ms.value / 1e9:7.1f} TFLOP/s(1x) | producer wait-empty {m[0] / 1e3:7.1f}k of {m[1] / 1e3:7.1f}k"
            f" | mma wait-full {m[2] / 1e3:7.1f}k wait-tmem {m[3] / 1e3:7.1f}k of {m[4] / 1e3:7.1f}k | epi wait-tfull {m[5] / 1e3:7.1f}k of {m[6] / 1e3:7.1f}k"
            f" [arrive {m[7] / 1e3:.1f}k tmem-ld {m[8] / 1e3:.1f}k math+st {m[9] / 1e3:.1f}k bar1 {m[10] / 1e3:.1f}k copy+bar2 {m[11] / 1e3:.1f}k]")


def main():
    torch.zeros(1, device="cuda")
    R = 50000
    for (K, N) in [(512, 512), (64, 512), (256, 256)]:
        for flags, name in [(0, "no output"), (11, "L0 relu"), (11 | 512, "L0 relu generic"), (26, "l1 relu"), (86, "l2 last"), (67, "dgrad out"),
                            (98, "dgrad l2 relu"), (103, "dgrad l1 relu"), (1 | 2 | 16 | 128, "l1 mish"), (2 | 64 | 256, "dgrad l2 mish"),
                            (1 | 2 | 4 | 64 | 256, "dgrad l1 mish")]:
            print(f"K={K} N={N} {name:16s}", run(R, K, N, flags), flush=True)
    print("-- tile width / ring depth, K=512 N=512, op images")
    for ntile, stages in [(256, 2), (256, 1), (128, 3), (128, 2), (64, 4), (64, 2)]:
        print(f"NTILE={ntile} stages={stages}", run(R, 512, 512, 26, ntile, stages), flush=True)
    print("-- no output")
    for ntile, stages in [(256, 2), (128, 3), (64, 4)]:
        print(f"NTILE={ntile} stages={stages}", run(R, 512, 512, 0, ntile, stages), flush=True)
    print("-- rows")
    for R2 in [6250, 12500, 25000, 100000]:
        print(f"R={R2}", run(R2, 512, 512, 26), flush=True)


if __name__ == "__main__":
    main()
