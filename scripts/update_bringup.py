"""Bring-up of the tensor-core update path on a B200: every kernel against float64 torch, errors printed (no asserts),
including a sweep of MN-major descriptor encodings for the wgrad kernel.  Usage: python scripts/update_bringup.py"""

import ctypes as C
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dppo_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = "cuda:0"
g = torch.Generator(device="cpu").manual_seed(0)


def rnd(*s):
    return torch.randn(*s, generator=g).to(dev)


def mish(x):
    return x * torch.tanh(torch.nn.functional.softplus(x))


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def linear_case(R, K, N, transposed=False, bias=True, pre=0, res=False, act_out=0):
    x = rnd(R, K)
    W = rnd(K, N) if transposed else rnd(N, K)
    b = rnd(N) if bias else None
    p = rnd(R, N) if pre else None
    r = rnd(R, N) if res else None
    out = torch.full((R, N), float("nan"), device=dev)
    oact = torch.full((R, N), float("nan"), device=dev)
    rc = lib.dppo_debug_linear(_lib.ptr(x), R, K, _lib.ptr(W), N, int(transposed), _lib.ptr(b), _lib.ptr(p), pre, _lib.ptr(r),
                               _lib.ptr(out), act_out, _lib.ptr(oact), _lib.stream_ptr())
    torch.cuda.synchronize()
    if rc != 0:
        return f"rc={rc} {lib.dppo_last_error().decode()}"
    Wd = (W.t() if transposed else W).double()
    ref = x.double() @ Wd.t()
    if b is not None:
        ref = ref + b.double()
    if pre:
        pd = p.double().requires_grad_(True)
        a = torch.relu(pd) if pre == 1 else mish(pd)
        (gr,) = torch.autograd.grad(a.sum(), pd)
        ref = ref * gr
    if r is not None:
        ref = ref + r.double()
    ra = ref if act_out == 0 else (torch.relu(ref) if act_out == 1 else mish(ref))
    return f"out {relerr(out, ref):.2e}  act_out {relerr(oact, ra):.2e}"


def wgrad_case(R, N, K, with_bias=True):
    gm, x = rnd(R, N), rnd(R, K)
    dW = torch.zeros(N, K, device=dev)
    db = torch.zeros(N, device=dev)
    rc = lib.dppo_debug_wgrad(_lib.ptr(gm), _lib.ptr(x), R, N, K, _lib.ptr(dW), _lib.ptr(db) if with_bias else None, _lib.stream_ptr())
    torch.cuda.synchronize()
    if rc != 0:
        return f"rc={rc} {lib.dppo_last_error().decode()}", 1.0
    ref = gm.double().t() @ x.double()
    e = relerr(dW, ref)
    eb = relerr(db, gm.double().sum(0)) if with_bias else 0.0
    return f"dW {e:.2e}  db {eb:.2e}", max(e, eb)


def main():
    print("== row GEMM (forward / dgrad form)")
    cases = [
        dict(R=300, K=64, N=512), dict(R=1000, K=512, N=512, act_out=1), dict(R=257, K=512, N=24, act_out=0),
        dict(R=130, K=256, N=1), dict(R=500, K=192, N=1024, act_out=2), dict(R=640, K=24, N=512, transposed=True, bias=False),
        dict(R=700, K=512, N=512, transposed=True, bias=False, pre=1, res=True), dict(R=333, K=512, N=512, transposed=True, bias=False, pre=2),
        dict(R=20000, K=512, N=512, act_out=1, res=True), dict(R=128, K=17, N=256, act_out=2), dict(R=999, K=1024, N=80),
        dict(R=450, K=512, N=64, bias=True),
    ]
    for c in cases:
        try:
            print(c, "->", linear_case(**c), flush=True)
        except Exception as ex:  # keep going: this is a survey
            print(c, "-> EXC", ex, flush=True)
            return
    print("== wgrad (MN-major operands)")
    wcases = [(300, 512, 512), (1000, 24, 512), (777, 512, 64), (130, 1, 256), (600, 1024, 192), (50, 256, 17), (20000, 512, 512),
              (64, 128, 64), (8192, 256, 256)]
    msg, err = wgrad_case(64, 128, 64)
    print("default descriptor, (64,128,64):", msg, flush=True)
    if err > 1e-3:
        print("-- default MN-major descriptor is wrong; sweeping (LBO, SBO)")
        for lbo, sbo in itertools.product([8192, 1024, 128, 2048, 16384, 0, 16], [1024, 8192, 128, 2048, 64, 16]):
            lib.dppo_debug_set_mn_desc(lbo, sbo)
            try:
                msg, err = wgrad_case(64, 128, 64, with_bias=False)
            except Exception as ex:
                print("EXC", ex)
                return
            print(f"   LBO={lbo} SBO={sbo}: {msg}", flush=True)
            if err < 1e-3:
                print("   ^^ works")
                break
    for c in wcases:
        try:
            print(c, "->", wgrad_case(*c)[0], flush=True)
        except Exception as ex:
            print(c, "-> EXC", ex, flush=True)
            return
    print("== full update vs torch autograd")
    from dppo_b200.workloads import get_workload
    from tests.helpers import GOLDEN_CASES, build_model, make_inputs, our_classes

    for case in ["hopper", "transport_k20", "furniture"]:
        spec = GOLDEN_CASES[case]
        w = get_workload(spec["workload"])
        E, ft = spec["n_envs"], w["ft_denoising_steps"]
        inp = make_inputs(w, E, 300 if case != "furniture" else 200)
        res = {}
        for mode in ("autograd", "fused"):
            os.environ["DPPO_B200_UPDATE"] = mode
            model = build_model(w, dev, our_classes())
            with torch.no_grad():
                out = model(cond={"state": inp["state"].to(dev)}, noise=inp["noise"].to(dev))
                lp = model.get_logprobs({"state": inp["state"].to(dev)}, out.chains).view(E, ft, w["horizon_steps"], w["action_dim"])
            b, d = inp["mb_b"].to(dev), inp["mb_d"].to(dev)
            state = inp["state"].to(dev)
            r = model.loss({"state": state[b]}, out.chains[b, d], out.chains[b, d + 1], d, inp["returns"].to(dev)[b],
                           inp["oldvalues"].to(dev)[b], inp["advantages"].to(dev)[b], lp[b, d] + inp["lp_shift"].to(dev),
                           reward_horizon=w["act_steps"])
            (r[0] + 0.5 * r[2]).backward()
            torch.cuda.synchronize()
            grads = {n: p.grad.detach().clone() for n, p in list(model.actor_ft.named_parameters()) + [("critic." + n, p) for n, p in model.critic.named_parameters()]}
            res[mode] = ([float(r[0]), float(r[2]), r[3], r[4], r[5]], grads)
        print(case, "scalars autograd", res["autograd"][0])
        print(case, "scalars fused   ", res["fused"][0])
        worst = []
        for n, ga in res["autograd"][1].items():
            gf = res["fused"][1][n]
            worst.append((relerr(gf, ga), n))
        worst.sort(reverse=True)
        print(case, "worst gradient rel-err (of max-norm):", [(f"{e:.2e}", n) for e, n in worst[:6]], flush=True)


if __name__ == "__main__":
    main()
