"""
In-tree build of libdppo_b200.so (hand-written sm_100a CUDA behind the C ABI of include/dppo_b200.h).

    python -m dppo_b200.build            # rebuilds when a source is newer than the library

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the working tree to the GPU box.
"""

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib", "libdppo_b200.so")
SOURCES = ["api.cu", "pack.cu", "chain_mlp.cu", "chain_pair.cu", "chain_small.cu", "unet_plan.cu", "chain_unet.cu", "update.cu", "update_gemm.cu", "update_plan.cu"]
# bring-up code (descriptor self-test, micro-benchmarks) lives in its own test-only library, not in the product .so
TEST_LIB = os.path.join(PKG, "lib", "libdppo_b200_test.so")
TEST_SOURCES = ["umma_selftest.cu", "microbench.cu"]
HEADERS = ["common.cuh", "internal.h", "chain_mlp.cuh", "unet_plan.h", "update_gemm.h", os.path.join("..", "..", "include", "dppo_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def _stale():
    if not os.path.exists(LIB) or not os.path.exists(TEST_LIB):
        return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(TEST_LIB))
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + TEST_SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile the library if it is missing or older than its sources; returns the path.
    One `nvcc -c` per translation unit, run concurrently, then one link step."""
    if not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(os.path.dirname(LIB), "obj")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        prof = ["-DDPPO_CHAIN_PROF"] if os.environ.get("DPPO_B200_CHAIN_PROF") == "1" else []  # MMA-warp wait counters
        prof += os.environ.get("DPPO_B200_NVCC_DEFS", "").split()  # bring-up: extra -D switches for A/B variants
        cmd = [nvcc] + NVCC_FLAGS + prof + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return obj, res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES) + len(TEST_SOURCES), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_one, SOURCES + TEST_SOURCES))
    for objs, out in ((results[:len(SOURCES)], LIB), (results[len(SOURCES):], TEST_LIB)):
        link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"] + [o for o, _ in objs] + ["-o", out]
        res = subprocess.run(link, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc link failed:\n" + " ".join(link) + "\n" + res.stdout + res.stderr)
    if verbose:
        print("".join(out for _, out in results) + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
