"""The C-ABI library builds, loads and exports every symbol include/dppo_b200.h declares (no compute, CPU only)."""

import ctypes
import os
import re

from dppo_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dppo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dppo_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(_lib.EXPORTS)


def test_library_exports_every_symbol():
    lib = _lib.load()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.dppo_version() >= 100


def test_errors_are_reported_not_raised_across_the_boundary():
    lib = _lib.load()
    rc = lib.dppo_ctx_create(None, None, None, 0, 0)
    assert rc == -1
    assert b"null" in lib.dppo_last_error()
    rc = lib.dppo_gae_f64(None, None, None, None, 4, 4, 0.99, 0.95, 1.0, None, None, None)
    assert rc == -1


def test_struct_layouts_match_header():
    # 10 int32
    assert ctypes.sizeof(_lib.MlpDesc) == 40
    # 4 int32 + 6 float + 9 pointers
    assert ctypes.sizeof(_lib.SchedDesc) == 40 + 9 * 8
    assert ctypes.sizeof(_lib.LossHp) == 5 * 4 + 7 * 4


def test_no_cpu_fallback():
    import pytest
    import torch

    with pytest.raises(RuntimeError):
        _lib.ptr(torch.zeros(4))
