"""
Host logic of the host-buffer rollout call without a GPU: VPGDiffusion.forward(cond={"state": host tensor}) ->
ChainEngine.sample_host -> dppo_sample_chain_host, with the C entry point replaced by a gcc-built stub that records its
arguments and fills the result buffers.  Checks the marshalling (pointers, flags, Philox offset sequence, deterministic /
base-policy / no-chain flags), the ring of result buffers, the dtype / layout conversion of the observations and the
weight-cache fast path.  (The real call is held to the device call bit for bit in tests/test_gpu_parity.py.)
"""

import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from dppo_b200 import _lib
from dppo_b200 import engine as E_
from dppo_b200.workloads import get_workload
from tests.helpers import build_model, our_classes

STUB = r"""
#include <stdint.h>
#include <string.h>
typedef struct { const float* state; int n; uint64_t seed, offset; int64_t env_offset; int det, base; float min_std;
                 float* traj; float* chain; int flags; void* stream; float state0; } call_t;
static call_t last; static int calls;
int D = 0, FT1 = 0, fail_next = 0;
int dppo_sample_chain_host(void* ctx, const float* state, int n, uint64_t seed, uint64_t offset, int64_t env_offset,
                           int det, int base, float min_std, float* traj, float* chain, int flags, void* stream) {
  if (fail_next) { fail_next = 0; return -2; }
  call_t c = {state, n, seed, offset, env_offset, det, base, min_std, traj, chain, flags, stream, state[0]};
  last = c; ++calls;
  for (int i = 0; i < n * D; ++i) traj[i] = (float)offset + 0.5f;
  if (chain) for (int i = 0; i < n * FT1 * D; ++i) chain[i] = (float)offset;
  return 0;
}
const call_t* stub_last(void) { return &last; }
int stub_calls(void) { return calls; }
const char* dppo_last_error(void) { return "stub failure"; }
"""


class Call(C.Structure):
    _fields_ = [("state", C.c_void_p), ("n", C.c_int), ("seed", C.c_uint64), ("offset", C.c_uint64), ("env_offset", C.c_int64),
                ("det", C.c_int), ("base", C.c_int), ("min_std", C.c_float), ("traj", C.c_void_p), ("chain", C.c_void_p),
                ("flags", C.c_int), ("stream", C.c_void_p), ("state0", C.c_float)]


@pytest.fixture()
def stubbed(tmp_path, monkeypatch):
    src = tmp_path / "stub.c"
    src.write_text(STUB)
    so = tmp_path / "libstub.so"
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", str(src), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    vp, i32, i64, u64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float
    lib.dppo_sample_chain_host.argtypes = [vp, vp, i32, u64, u64, i64, i32, i32, f32, vp, vp, i32, vp]
    lib.stub_last.restype = C.POINTER(Call)
    lib.dppo_last_error.restype = C.c_char_p
    w = get_workload("hopper")
    model = build_model(w, "cpu", our_classes())
    Ta, Da, ft = w["horizon_steps"], w["action_dim"], w["ft_denoising_steps"]
    C.c_int.in_dll(lib, "D").value = Ta * Da
    C.c_int.in_dll(lib, "FT1").value = ft + 1
    # an engine without a device context: only the host logic of sample_host / the weight-cache check runs
    eng = E_.ChainEngine.__new__(E_.ChainEngine)
    eng.lib, eng.ctx, eng.ft, eng.D, eng.S, eng.is_unet = lib, C.c_void_p(0x1234), ft, Ta * Da, 20, False
    eng.device, eng._packed, eng._plists, eng._host_io, eng._current = torch.device("cpu"), {0: None, 1: None}, {}, {}, None
    eng.Ta, eng.Da, eng.cond_numel = Ta, Da, w["obs_dim"] * w["cond_steps"]
    eng._sample_host_fn = lib.dppo_sample_chain_host
    packs = []

    def fake_sync(which, net):  # stands in for dppo_pack_mlp: records that a repack was requested
        ps = E_._mlp_param_list(net)
        eng._plists[which] = (net, ps)
        packs.append(which)

    eng.sync_weights = fake_sync

    def host_ring(E):  # pin_memory() needs CUDA: plain host tensors stand in for the page-locked ring
        ring = []
        for _ in range(eng.HOST_RING):
            t, c = torch.zeros((E, Ta, Da)), torch.zeros((E, ft + 1, Ta, Da))
            ring.append((t, c, t.data_ptr(), c.data_ptr()))
        return [0, ring]

    eng._host_ring = host_ring
    monkeypatch.setattr(E_, "_RAW_STREAM", lambda dev: 0x77)
    monkeypatch.setattr(E_, "_CUR_DEVICE", lambda: 0)
    monkeypatch.setattr(_lib, "load", lambda *a, **k: lib)
    model._engine = eng
    return w, model, eng, lib, packs


def test_host_call_marshalling_ring_and_flags(stubbed):
    w, model, eng, lib, packs = stubbed
    E, Ta, Da, ft = 7, w["horizon_steps"], w["action_dim"], w["ft_denoising_steps"]
    torch.manual_seed(1234)
    state = torch.rand(E, 1, w["obs_dim"])
    outs = []
    for i in range(6):
        out = model(cond={"state": state}, env_offset=5)
        c = lib.stub_last().contents
        assert (c.n, c.offset, c.env_offset, c.det, c.base, c.flags) == (E, i + 1, 5, 0, 0, _lib.HOST_OUT_PINNED)
        assert c.seed == 1234 and c.stream == 0x77 and c.state == state.data_ptr()  # small observations: staged by the library
        assert abs(c.min_std - float(model.get_min_sampling_denoising_std())) < 1e-7
        assert out.trajectories.shape == (E, Ta, Da) and out.chains.shape == (E, ft + 1, Ta, Da)
        assert c.traj == out.trajectories.data_ptr() and c.chain == out.chains.data_ptr()
        assert float(out.chains[0, 0, 0, 0]) == i + 1 and float(out.trajectories[-1, -1, -1]) == i + 1.5
        outs.append(out)
    # ring of four: the three results before the last one are intact, older buffers have been reused
    for i in (3, 4, 5):
        assert float(outs[i].chains[0, 0, 0, 0]) == i + 1
    assert outs[1].chains.data_ptr() == outs[5].chains.data_ptr() and outs[0].chains.data_ptr() != outs[5].chains.data_ptr()
    # flags / no chain
    out = model(cond={"state": state}, deterministic=True, use_base_policy=True, return_chain=False)
    c = lib.stub_last().contents
    assert (c.det, c.base, c.chain, out.chains) == (1, 1, None, None)
    # float64 / non-contiguous observations are converted, not reinterpreted
    st64 = torch.rand(E, 1, 2 * w["obs_dim"], dtype=torch.float64)[:, :, ::2]
    model(cond={"state": st64})
    c = lib.stub_last().contents
    assert c.state != st64.data_ptr() and abs(c.state0 - float(st64[0, 0, 0])) < 1e-6
    # wrong observation width / library error
    with pytest.raises(RuntimeError, match="expected"):
        model(cond={"state": torch.rand(E, 1, w["obs_dim"] + 1)})
    C.c_int.in_dll(lib, "fail_next").value = 1
    with pytest.raises(RuntimeError, match="stub failure"):
        model(cond={"state": state})
    # a batch size of its own ring; E = 0 launches nothing
    n = lib.stub_calls()
    empty = model(cond={"state": torch.zeros(0, 1, w["obs_dim"])})
    assert empty.trajectories.shape == (0, Ta, Da) and lib.stub_calls() == n


def test_weight_cache_fast_path_sees_every_kind_of_change(stubbed):
    w, model, eng, lib, packs = stubbed
    state = torch.rand(3, 1, w["obs_dim"])
    model(cond={"state": state})
    assert packs == [0, 1]  # first decision: both networks packed
    model(cond={"state": state})
    assert packs == [0, 1]  # unchanged weights: the combined signature matches, no repack
    with torch.no_grad():
        next(model.actor_ft.parameters()).add_(1e-3)  # an optimiser step bumps the version counter
    model(cond={"state": state})
    assert packs == [0, 1, 0, 1]
    with torch.no_grad():
        list(model.actor.parameters())[-1].mul_(1.0)  # the frozen base policy too (load_state_dict)
    model(cond={"state": state})
    assert len(packs) == 6
    model.step()  # no annealing configured: nothing changes
    model(cond={"state": state})
    assert len(packs) == 6
    import copy

    model.actor_ft = copy.deepcopy(model.actor_ft)  # a replaced module (annealing swaps actor <- actor_ft)
    model(cond={"state": state})
    assert len(packs) == 8
    # a registered forward hook takes nn.Module.__call__'s own path and still reaches forward
    seen = []
    h = model.register_forward_hook(lambda m, a, o: seen.append(1))
    model(cond={"state": state})
    h.remove()
    assert seen == [1]
