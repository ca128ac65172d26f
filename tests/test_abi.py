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


def test_header_is_self_contained_c_and_a_c_program_links(tmp_path):
    """include/dppo_b200.h compiles on its own as C99 and as C++ (a cgo / JNI / plain C consumer includes nothing else), and
    a C program that includes it links against libdppo_b200.so and gets errors back as codes (no GPU needed)."""
    import subprocess

    prog = tmp_path / "use.c"
    prog.write_text(
        '#include "dppo_b200.h"\n#include <stdio.h>\n'
        "int main(void) {\n"
        "  float obs[4] = {0}, traj[4];\n"
        "  int rc = dppo_sample_chain_host(0, obs, 1, 0, 0, 0, 0, 0, 0.1f, traj, 0, DPPO_HOST_STATE_PINNED | DPPO_HOST_OUT_PINNED, 0);\n"
        '  printf("%d %d %s\\n", dppo_version(), rc, dppo_last_error());\n'
        "  return rc == DPPO_ERR_INVALID ? 0 : 1;\n}\n")
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", str(prog)], check=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", inc, "-x", "c++", "-fsyntax-only", str(prog)], check=True)
    _lib.load()
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = tmp_path / "use"
    subprocess.run(["gcc", "-std=c99", "-I", inc, str(prog), "-L", libdir, "-ldppo_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    version, rc, msg = out.stdout.split(" ", 2)
    assert int(version) >= 103 and int(rc) == -1 and "null" in msg
