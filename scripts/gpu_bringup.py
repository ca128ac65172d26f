"""
GPU bring-up: runs every kernel family in its own subprocess (a trap in one must not poison the others), with a
timeout, and writes gpurun_out/bringup.log.  Usage on the GPU box:  python scripts/gpu_bringup.py [stage ...]
"""

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")

STAGES = {
    "selftest": """
import torch
from dppo_b200 import _lib
lib = _lib.load_test()
g = torch.Generator().manual_seed(0)
for N, K in [(32, 64), (64, 64), (64, 128), (64, 256)]:
    a = torch.randn(128, K, generator=g).cuda(); b = torch.randn(N, K, generator=g).cuda()
    c = torch.zeros(128, N, device='cuda'); scratch = torch.zeros(K // 64 * 16384, dtype=torch.uint8, device='cuda')
    rc = lib.dppo_selftest_umma(_lib.ptr(a), _lib.ptr(b), _lib.ptr(c), _lib.ptr(scratch), N, K, 0, 0, _lib.stream_ptr())
    torch.cuda.synchronize()
    ref = a.bfloat16().float() @ b.bfloat16().float().T
    print('selftest', N, K, 'rc', rc, 'maxerr', float((c - ref).abs().max()), 'refmax', float(ref.abs().max()))
""",
    "gae": """
import pytest, sys
sys.exit(pytest.main(['-x', '-q', 'tests/test_gpu_parity.py::test_gae_matches_oracle']))
""",
    "chain_hopper": """
import numpy as np, torch
from tests.helpers import *
from tests.test_gpu_parity import rel_err
import time
for case in __CASES__:
    spec = GOLDEN_CASES[case]; w = get_workload(spec['workload'])
    model = build_model(w, 'cuda:0', our_classes()); gold = load_golden(case)
    inp = make_inputs(w, spec['n_envs'], spec['mb_rows'])
    t0 = time.time()
    out = model(cond={'state': inp['state'].cuda()}, noise=inp['noise'].cuda())
    torch.cuda.synchronize()
    e = rel_err(out.chains.cpu().numpy(), gold['chains'])
    print(case, 'chains maxrel', e.max(), 'frac>1e-3', (e > 1e-3).mean(), 'first call s', round(time.time() - t0, 3))
    for s in range(e.shape[1]):
        print('   slot', s, 'maxrel', e[:, s].max())
    with torch.no_grad():
        lp = model.get_logprobs({'state': inp['state'].cuda()}, torch.from_numpy(gold['chains']).cuda())
    torch.cuda.synchronize()
    e2 = rel_err(lp.cpu().numpy(), gold['logprobs'])
    print(case, 'logprobs maxrel', e2.max(), 'frac>1e-3', (e2 > 1e-3).mean())
""",
}
STAGES["chain_all"] = STAGES["chain_hopper"].replace("__CASES__", "['walker2d', 'transport_k20', 'transport', 'furniture', 'furniture_ddpm100']")
STAGES["chain_hopper"] = STAGES["chain_hopper"].replace("__CASES__", "['hopper']")
STAGES["loss"] = """
import pytest, sys
sys.exit(pytest.main(['-x', '-q', 'tests/test_gpu_parity.py', '-k', 'loss']))
"""


def main():
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or list(STAGES)
    with open(os.path.join(OUT, "bringup.log"), "a") as log:
        for name in names:
            t0 = time.time()
            try:
                res = subprocess.run([sys.executable, "-c", STAGES[name]], cwd=ROOT, capture_output=True, text=True, timeout=300)
                out, rc = res.stdout + res.stderr, res.returncode
            except subprocess.TimeoutExpired as e:
                out, rc = (e.stdout or b"").decode() + (e.stderr or b"").decode() + "\nTIMEOUT", -999
            msg = f"===== {name}: rc={rc} ({time.time() - t0:.1f}s)\n{out[-6000:]}\n"
            print(msg)
            log.write(msg)
            log.flush()


if __name__ == "__main__":
    main()
