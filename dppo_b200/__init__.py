"""
dppo_b200 — B200-native (sm_100a) implementation of the DPPO data-parallel hot path.

Host side (this package, Python/PyTorch) mirrors the reference's model/agent interface for that path; the arithmetic
runs in hand-written CUDA kernels behind a C-ABI shared library (include/dppo_b200.h, dppo_b200/csrc/).
"""

__version__ = "0.1.0"
